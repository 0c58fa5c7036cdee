// Tail kernel with one CTA per cutout (PM_PATH_SPLIT; PM_TAIL_WARP selects one warp per cutout instead, pm_tail.cu): r20/r80, calcA x3,
// smoothness (tail_one).
// PM_TAILCTA_THREADS (tuning): a smaller group (e.g. 128 threads, more CTAs per SM) in its own namespace
#ifdef PM_TAILCTA_THREADS
#define PM_THREADS PM_TAILCTA_THREADS
#define PM_NS pmk_t
#endif
#include "pm_launch.cuh"
#ifndef PM_TAILCTA_MIN_CTAS
#define PM_TAILCTA_MIN_CTAS PM_MIN_CTAS
#endif
PM_DEFINE_KERNEL(pm_tail_cta, tail_body, PM_NS::kThreads, PM_TAILCTA_MIN_CTAS)
extern "C" int pm_tail_cta_min_ctas(void) { return PM_TAILCTA_MIN_CTAS; }
extern "C" int pm_tail_cta_threads(void) { return PM_NS::kThreads; }
