// pm_tail.cu — the tail kernel of the split path with ONE WARP PER CUTOUT (casgm.calcR20_R80 / concentration,
// asymmetry.calcA x3, casgm.smoothness; main.py:470-495).
//
// These phases are chains of dependent operations (FP64 square-root chains of the circle/pixel overlap, the
// root finder, gathers through the bit planes): with a whole CTA per cutout most warps sit at barriers or idle
// while one of them walks the chain.  Compiled with PM_GROUP_WARP the same phase code (pm_kernels.cuh) runs on a
// group of 32 threads with __syncwarp as its barrier and warp shuffles as its reductions, so that 16 (or 32)
// independent cutouts are in flight per SM instead of two.  Each group owns an equal slice of the CTA's dynamic
// shared memory (the segmap bit plane, the row offsets, block scalars); what does not fit lives in the group's
// slice of the workspace (L1/L2 resident).
#define PM_GROUP_WARP 1
#define PM_NS pmk_w
#include "pm_launch.cuh"

#ifndef PM_TAIL_GROUPS_PER_CTA
#define PM_TAIL_GROUPS_PER_CTA 4
#endif
#ifndef PM_TAIL_CTAS_PER_SM
#define PM_TAIL_CTAS_PER_SM 4
#endif
PM_DEFINE_KERNEL(pm_tail_warp, tail_body, 32 * PM_TAIL_GROUPS_PER_CTA, PM_TAIL_CTAS_PER_SM)

extern "C" {
int pm_tail_warp_groups_per_cta(void) { return PM_TAIL_GROUPS_PER_CTA; }
int pm_tail_warp_ctas_per_sm(void) { return PM_TAIL_CTAS_PER_SM; }
// shared memory a group cannot do without (the block scalars); planes are placed while room is left
unsigned pm_tail_warp_fixed_smem(int n) { return pmk_w::smem_fixed_bytes(n, 1, 0); }
}
