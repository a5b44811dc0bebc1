// agb_kernels_p3b.cu — instantiates the instance kernels for 3 player(s), big layout (one TU each: parallel builds).
#include "agb_kernels.cuh"
namespace agb {
cudaError_t set_attr_p3b(int model, size_t smem) { return set_attr_p<3, true>(model, smem); }
void launch_solve_p3b(const LaunchArgs& L) { launch_solve_p<3, true>(L); }
void launch_op_p3b(const LaunchArgs& L) { launch_op_p<3, true>(L); }
void launch_ibr_p3b(const LaunchArgs& L) { launch_ibr_p<3, true>(L); }
}  // namespace agb
