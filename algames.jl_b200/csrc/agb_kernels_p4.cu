// agb_kernels_p4.cu — instantiates the instance kernels for 4 player(s), layout 1 (see agb_kernels.cuh; one TU each: parallel builds).
#include "agb_kernels.cuh"
namespace agb {
cudaError_t set_attr_p4(int model, size_t smem) { return set_attr_p<4, 1>(model, smem); }
void launch_solve_p4(const LaunchArgs& L) { launch_solve_p<4, 1>(L); }
void launch_op_p4(const LaunchArgs& L) { launch_op_p<4, 1>(L); }
void launch_ibr_p4(const LaunchArgs& L) { launch_ibr_p<4, 1>(L); }
}  // namespace agb
