// Front kernel of the split path, one CTA per cutout: pixmap.pixelmap / checkPixelmapEdges / calcRmax, aperpixmap, the
// sort, casgm.gini / m20 and the centroid candidates of asymmetry.minapix (front_one).  The only pass that streams the
// image from HBM (TMA-staged bands).
#include "pm_launch.cuh"
PM_DEFINE_KERNEL(pm_front, front_body, PM_NS::kThreads, PM_MIN_CTAS)
