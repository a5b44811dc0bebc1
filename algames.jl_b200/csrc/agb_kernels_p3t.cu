// agb_kernels_p3t.cu — instantiates the instance kernels for 3 player(s), layout 3 (see agb_kernels.cuh; one TU each: parallel builds).
#include "agb_kernels.cuh"
namespace agb {
cudaError_t set_attr_p3t(int model, size_t smem) { return set_attr_p<3, 3>(model, smem); }
void launch_solve_p3t(const LaunchArgs& L) { launch_solve_p<3, 3>(L); }
void launch_op_p3t(const LaunchArgs& L) { launch_op_p<3, 3>(L); }
void launch_ibr_p3t(const LaunchArgs& L) { launch_ibr_p<3, 3>(L); }
}  // namespace agb
