// agb_kernels_p2.cu — instantiates the instance kernels for 2 player(s), layout 0 (see agb_kernels.cuh; one TU each: parallel builds).
#include "agb_kernels.cuh"
namespace agb {
cudaError_t set_attr_p2(int model, size_t smem) { return set_attr_p<2, 0>(model, smem); }
void launch_solve_p2(const LaunchArgs& L) { launch_solve_p<2, 0>(L); }
void launch_op_p2(const LaunchArgs& L) { launch_op_p<2, 0>(L); }
void launch_ibr_p2(const LaunchArgs& L) { launch_ibr_p<2, 0>(L); }
}  // namespace agb
