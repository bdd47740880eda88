// Shared declarations of libfs2b200.so (internal; the public ABI is include/fs2_b200.h).
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/fs2_b200.h"

namespace fs2 {

void set_error(const char* fmt, ...);
extern unsigned long long g_kernel_launches;  // every kernel this library enqueues (fs2_kernel_launches())

#define FS2_CUDA_CHECK(expr)                                                                   \
  do {                                                                                         \
    cudaError_t _e = (expr);                                                                   \
    if (_e != cudaSuccess) {                                                                   \
      fs2::set_error("%s failed at %s:%d: %s", #expr, __FILE__, __LINE__, cudaGetErrorString(_e)); \
      return FS2_ERR_CUDA;                                                                     \
    }                                                                                          \
  } while (0)

#define FS2_LAUNCH_CHECK()                 \
  do {                                     \
    ++fs2::g_kernel_launches;              \
    FS2_CUDA_CHECK(cudaGetLastError());    \
  } while (0)

#define FS2_REQUIRE(cond, ...)      \
  do {                              \
    if (!(cond)) {                  \
      fs2::set_error(__VA_ARGS__);  \
      return FS2_ERR_INVALID;       \
    }                               \
  } while (0)

// ---- operator launchers (one per kernel family) -------------------------------------------
// All of them enqueue on `st` and return FS2_OK / error code.

enum Act { ACT_NONE = 0, ACT_RELU = 1, ACT_TANH = 2 };

// out[m, n] = act( sum_{j<taps} sum_{k<K} x[b, t+j-pad, k] * w[j][n][k] + bias[n] ) (+ resid[m, n])
// x rows have stride ldx floats, time extent L per utterance (zero outside [0,L)); m = b*L + t.
struct TapGemm {
  const float* x; int ldx;
  int B, L, K;
  const float* w;      // [taps][N][K]
  const float* bias;   // [N] or nullptr
  int N, taps;
  int act;
  const float* resid; int ldr;  // nullptr => none
  float* out; int ldo;
  // tensor-core families, N == 384, taps == 1: fuse LayerNorm over the full output row into the epilogue (gemm_ln_tc.cu)
  const float* ln_gamma = nullptr; const float* ln_beta = nullptr; float ln_eps = 0.f;
  // tensor-core families only: store output columns >= vt_col0 transposed into vt_out (see gemm_tc.cu)
  float* vt_out = nullptr; int vt_col0 = 0, vt_dk = 0, vt_heads = 0, vt_lpad = 0;
  // f16 family (tap_gemm_f16): fp16 copies of the activations (row stride ldx_h halfs) and of w; the result goes to
  // out (fp32, with the optional residual) and / or out_h (fp16, row stride ldo_h halfs, for the next f16 GEMM)
  const __half* x_h = nullptr; int ldx_h = 0; const __half* w_h = nullptr; __half* out_h = nullptr; int ldo_h = 0;
  // 3xF16 (the error-compensated family): w split into fp16 hi = rn(w), lo = rn(w - hi), same [taps][N][K] layout
  const __half* w_hi_h = nullptr; const __half* w_lo_h = nullptr;
  __half* split_ws = nullptr;   // scratch for the fp16 hi / lo planes of x: 2 * B*L * K halfs (= the bytes of x)
  bool split_ready = false;     // the producer of x already wrote the planes (row_norm's split_out): skip the pre-pass
};
int tap_gemm_fp32(const TapGemm& g, cudaStream_t st);
int tap_gemm_tf32(const TapGemm& g, cudaStream_t st);   // tcgen05 + TMA (gemm_tc.cu)
bool gemm_ln_tf32_supported(const TapGemm& g);         // row-complete GEMM + residual + LayerNorm (gemm_ln_tc.cu)
int gemm_ln_tf32(const TapGemm& g, cudaStream_t st);
int tap_gemm_3xtf32(const TapGemm& g, cudaStream_t st); // same kernel, error-compensated split operands ("3xF16")
int tap_gemm_f16(const TapGemm& g, cudaStream_t st);    // same kernel, kind::f16 on x_h / w_h, fp32 accumulation
int to_half(const float* src, __half* dst, long n, cudaStream_t st);
int split_f16(const float* src, __half* hi, __half* lo, long n, cudaStream_t st);   // round to nearest, clamped to +-65504
constexpr int MATH_3XTF32 = FS2_MATH_3XTF32;            // also what the other tensor-core modes use for the encoder + predictors

// Row LayerNorm with the fusions the path needs.
struct RowNorm {
  const float* x; int ldx;        // [rows, C]
  const float* resid; int ldr;    // optional, added before the statistics
  const float* gamma; const float* beta; float eps;
  int64_t rows; int C;            // C in {256, 384}
  float* out; int ldo;            // optional (nullptr when only the head is wanted)
  int relu_after;                 // y = relu(LN(.))   (decoder input layer, core/encoder.py:118-125)
  const float* pe; const float* alpha; int L;  // optional: y += alpha * pe[t], t = row % L
  // optional scalar head (predictors): s = y . head_w + head_b, 0 where t >= lens[b]
  const float* head_w; const float* head_b; float* head_out; int64_t* dur_out;
  const int64_t* lens;            // optional mask for the head outputs
  __half* out_h; int ldo_h;       // optional fp16 copy of y (A operand of an f16 GEMM)
  __half* split_out;              // optional 3xF16 planes of y: hi at [row][C], lo at [rows + row][C] (A operand of a 3xF16 GEMM)
};
int row_norm(const RowNorm& r, cudaStream_t st);

int embed_posenc(const int64_t* xs, const float* table, int n_sym, const float* pe, const float* alpha, int B, int T,
                 int C, float* out, cudaStream_t st);

int bucketize(const float* vals, const float* bins, int n_edges, int64_t n, int64_t* ids, cudaStream_t st);
int one_hot(const int64_t* ids, int64_t n, int n_bins, float* out, cudaStream_t st);
// out[r,:] = (hm[r,:] + (p_tab[p_id[r]] + p_bias)) + (e_tab[e_id[r]] + e_bias); ids from values (nullable id outs)
int variance_embed_add(const float* hm, const float* e_val, const float* p_val, const float* e_bins, const float* p_bins,
                       int n_edges, const float* e_tab, const float* e_bias, const float* p_tab, const float* p_bias,
                       int64_t rows, int C, float* out, int64_t* e_ids, int64_t* p_ids, cudaStream_t st);

int attention_fp32(const float* qkv, const int64_t* lens, int B, int L, int C, int heads, float* ctx, cudaStream_t st);
// tensor-core attention: q, k from qkv [B,L,3C]; v from the transposed buffer vt [B*heads, dk, lpad]
int attention_tf32(const float* qkv, const float* vt, int lpad, const int64_t* lens, int B, int L, int C, int heads,
                   float* ctx, cudaStream_t st);

// vt[(b*heads + h)*dk + d][t] = qkv[b, t, 2C + h*dk + d]   (test helper for the single-operator entry)
int transpose_v(const float* qkv, int B, int L, int C, int heads, float* vt, int lpad, cudaStream_t st);

int length_plan(void* ds, int ds_dtype, const int64_t* ilens, float alpha, int B, int T, int mutate, int32_t* cum,
                int64_t* olens, int64_t* stats, cudaStream_t st);
int length_gather(const float* hs, const int32_t* cum, const int64_t* ilens, int B, int T, int C, float* out, int Lcap,
                  cudaStream_t st);

int masked_losses(const float* before, const float* after, const float* ys, int ld_ys_time, const float* d_out,
                  const void* ds, int ds_dtype, const float* e_out, const float* p_out, const float* es, const float* ps,
                  const int64_t* ilens, const int64_t* olens, int B, int T, int L, int odim, float* out7, void* scratch,
                  cudaStream_t st);

// weight repacking helpers (pack.cu)
int pack_conv_weight(const float* src, int N, int K, int taps, const float* scale /*[N] or null*/, float* dst,
                     cudaStream_t st);  // src [N][K][taps] -> dst [taps][N][K] * scale[n]
int pack_transpose(const float* src, int rows, int cols, float* dst, cudaStream_t st);  // [rows][cols] -> [cols][rows]
int fold_batchnorm(const float* gamma, const float* beta, const float* mean, const float* var, float eps, int N,
                   float* scale, float* shift, cudaStream_t st);

// ---- small device helpers -----------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
// torch.bucketize(right=False): first i with !(bins[i] < x) ... written as torch does so NaN -> n_edges
__device__ __forceinline__ int bucket_of(float x, const float* __restrict__ bins, int n_edges) {
  int lo = 0, hi = n_edges;
  while (lo < hi) {
    int mid = lo + ((hi - lo) >> 1);
    if (!(bins[mid] >= x)) lo = mid + 1; else hi = mid;
  }
  return lo;
}

}  // namespace fs2
