// Interim: until attention_tc.cu lands, FS2_MATH_TF32 runs the attention core on the exact-fp32
// streaming-softmax kernel (more precise, slower).  The dense projections around it are tcgen05.
#include "common.cuh"
namespace fs2 {
int attention_tf32(const float* qkv, const int64_t* lens, int B, int L, int C, int heads, float* ctx, cudaStream_t st) {
  return attention_fp32(qkv, lens, B, L, C, heads, ctx, st);
}
}
