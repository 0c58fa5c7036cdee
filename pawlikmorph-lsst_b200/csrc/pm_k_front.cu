// Front kernel of the split path, one CTA per cutout: pixmap.pixelmap / checkPixelmapEdges / calcRmax, aperpixmap, the
// sort, casgm.gini / m20 and the centroid candidates of asymmetry.minapix (front_one).  The only pass that streams the
// image from HBM (TMA-staged bands).
#ifdef PM_FRONT_THREADS
#define PM_THREADS PM_FRONT_THREADS
#define PM_NS pmk_f
#endif
#include "pm_launch.cuh"
#ifndef PM_FRONT_MIN_CTAS
#define PM_FRONT_MIN_CTAS PM_MIN_CTAS
#endif
PM_DEFINE_KERNEL(pm_front, front_body, PM_NS::kThreads, PM_FRONT_MIN_CTAS)
extern "C" int pm_front_min_ctas(void) { return PM_FRONT_MIN_CTAS; }
extern "C" int pm_front_threads(void) { return PM_NS::kThreads; }
