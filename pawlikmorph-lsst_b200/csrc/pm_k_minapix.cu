// asymmetry.minapix (asymmetry.py:113-129), one CTA per cutout: the loop needs ~64 registers and the window + pixel
// list in shared memory, so more CTAs fit an SM than in the front kernel (minapix_one).
#ifdef PM_MINAPIX_THREADS
#define PM_THREADS PM_MINAPIX_THREADS
#define PM_NS pmk_m
#endif
#include "pm_launch.cuh"
#ifndef PM_MINAPIX_MIN_CTAS
#define PM_MINAPIX_MIN_CTAS 4
#endif
PM_DEFINE_KERNEL(pm_minapix, minapix_body, PM_NS::kThreads, PM_MINAPIX_MIN_CTAS)
extern "C" int pm_minapix_min_ctas(void) { return PM_MINAPIX_MIN_CTAS; }
extern "C" int pm_minapix_threads(void) { return PM_NS::kThreads; }
