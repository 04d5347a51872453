// NTT kernels instantiated for one field (separate translation unit: ptxas time runs in parallel).
#include "ntt_host.cuh"
namespace plk { const NttOps* ntt_ops_bls12_377_base() { return make_ntt_ops<Bls12377BaseParams>(); } }
