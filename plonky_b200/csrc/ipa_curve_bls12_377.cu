// Halo IPA round kernels instantiated for one curve (separate translation unit: ptxas time runs in parallel).
#include "ipa_kernels.cuh"
namespace plk { const IpaOps* ipa_ops_bls12_377() { return make_ipa_ops<Bls12377Params>(); } }
