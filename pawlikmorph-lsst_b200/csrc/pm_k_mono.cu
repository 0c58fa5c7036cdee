// The monolith: one CTA per cutout through every phase (analyse_one).  The whole batch by default, and — with PM_PATH_SPLIT — the
// pass over the cutouts the front kernel of the split path hands over (lists / radii beyond the hand-over slots).
#include "pm_launch.cuh"
PM_DEFINE_KERNEL(pm_analyse, analyse_body, PM_NS::kThreads, PM_MIN_CTAS)
PM_DEFINE_KERNEL(pm_big, big_body, PM_NS::kThreads, PM_MIN_CTAS)
