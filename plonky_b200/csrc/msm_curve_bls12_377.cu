// MSM kernels instantiated for one curve (separate translation unit: ptxas time runs in parallel).
#include "msm_kernels.cuh"
namespace plk { const MsmOps* msm_ops_bls12_377() { return make_msm_ops<Bls12377Params>(); } }
