// MSM kernels instantiated for one curve (separate translation unit: ptxas time runs in parallel).
#include "msm_kernels.cuh"
namespace plk { const MsmOps* msm_ops_tweedledum() { return make_msm_ops<TweedledumParams>(); } }
