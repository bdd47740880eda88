// placeholder until gemm_tc.cu / attention_tc.cu land
#include "common.cuh"
namespace fs2 {
int tap_gemm_tf32(const TapGemm&, cudaStream_t) { set_error("tf32 tensor-core GEMM not built yet"); return FS2_ERR_INVALID; }
int attention_tf32(const float*, const int64_t*, int, int, int, int, float*, cudaStream_t) { set_error("tf32 tensor-core attention not built yet"); return FS2_ERR_INVALID; }
}
