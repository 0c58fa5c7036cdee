// Tail kernel with one CTA per cutout (PM_TAIL_CTA; the default is one warp per cutout, pm_tail.cu): r20/r80, calcA x3,
// smoothness (tail_one).
#include "pm_launch.cuh"
PM_DEFINE_KERNEL(pm_tail_cta, tail_body, PM_NS::kThreads, PM_MIN_CTAS)
