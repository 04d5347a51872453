// NTT kernels instantiated for one field (separate translation unit: ptxas time runs in parallel).
#include "ntt_host.cuh"
namespace plk { const NttOps* ntt_ops_tweedledee_base() { return make_ntt_ops<TweedledeeBaseParams>(); } }
