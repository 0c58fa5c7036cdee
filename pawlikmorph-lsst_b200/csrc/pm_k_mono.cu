// The monolith: one CTA per cutout through every phase (analyse_one).  The whole batch by default, and — with PM_PATH_SPLIT — the
// pass over the cutouts the front kernel of the split path hands over (lists / radii beyond the hand-over slots).
// PM_MONO_THREADS / PM_MONO_MIN_CTAS (tuning): another CTA shape, in its own namespace
#ifdef PM_MONO_THREADS
#define PM_THREADS PM_MONO_THREADS
#define PM_NS pmk_mono
#endif
#include "pm_launch.cuh"
#ifndef PM_MONO_MIN_CTAS
#define PM_MONO_MIN_CTAS PM_MIN_CTAS
#endif
PM_DEFINE_KERNEL(pm_analyse, analyse_body, PM_NS::kThreads, PM_MONO_MIN_CTAS)
PM_DEFINE_KERNEL(pm_big, big_body, PM_NS::kThreads, PM_MONO_MIN_CTAS)
extern "C" int pm_mono_threads(void) { return PM_NS::kThreads; }
extern "C" int pm_mono_min_ctas(void) { return PM_MONO_MIN_CTAS; }
