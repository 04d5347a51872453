// Hash-to-curve / point codec kernels instantiated for one curve (separate translation unit).
#include "codec_kernels.cuh"
namespace plk { const CodecOps* codec_ops_tweedledum() { return make_codec_ops<TweedledumParams>(); } }
