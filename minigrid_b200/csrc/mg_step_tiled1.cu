// mg_step_tiled1.cu — the MODE_TILED1 instantiations of K1 (tiled layout, one buffer per warp), see mg_step_kernel.cuh.
#include "mg_step_kernel.cuh"

namespace mg {

StepKernel step_kernel_tiled1(int kind, int vis) { return pick_vis<MODE_TILED1>(kind, vis); }

#ifdef MG_TIMELINE
int debug_timeline_tiled1(void *out) { return (int)cudaMemcpyFromSymbol(out, g_tl, sizeof(g_tl)); }
#endif

}  // namespace mg
