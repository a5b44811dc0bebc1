// agb_kernels_p3.cu — instantiates the instance kernels for 3 player(s), layout 0 (see agb_kernels.cuh; one TU each: parallel builds).
#include "agb_kernels.cuh"
namespace agb {
cudaError_t set_attr_p3(int model, size_t smem) { return set_attr_p<3, 0>(model, smem); }
void launch_solve_p3(const LaunchArgs& L) { launch_solve_p<3, 0>(L); }
void launch_op_p3(const LaunchArgs& L) { launch_op_p<3, 0>(L); }
void launch_ibr_p3(const LaunchArgs& L) { launch_ibr_p<3, 0>(L); }
}  // namespace agb
