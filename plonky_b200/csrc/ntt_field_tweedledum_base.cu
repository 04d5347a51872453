// NTT kernels instantiated for one field (separate translation unit: ptxas time runs in parallel).
#include "ntt_host.cuh"
namespace plk { const NttOps* ntt_ops_tweedledum_base() { return make_ntt_ops<TweedledumBaseParams>(); } }
