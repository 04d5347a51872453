// Hash-to-curve / point codec kernels instantiated for one curve (separate translation unit).
#include "codec_kernels.cuh"
namespace plk { const CodecOps* codec_ops_bls12_377() { return make_codec_ops<Bls12377Params>(); } }
