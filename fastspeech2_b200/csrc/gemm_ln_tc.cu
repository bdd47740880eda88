// Row-complete tensor-core GEMM with the FFT block's residual + LayerNorm fused into the epilogue:
//
//   out[m,:] = LayerNorm_C( x[m,:] . W^T + bias + resid[m,:] ) * gamma + beta          (C = 384)
//
// i.e. `x = norm(x + linear_out(ctx))` and `x = norm(x + w_2(hid))` of EncoderLayer.forward
// (core/encoder.py:60-69) as ONE kernel instead of GEMM -> y -> LayerNorm kernel.  These projections
// are HBM-bound (K <= 1024 against 3 x M x C x 4 bytes of compulsory traffic), so removing the write and
// re-read of y is what matters: 236 MB instead of 393 MB per out-projection at c2.
//
// A CTA owns 128 complete rows (all C columns): per 32-wide k-step two tcgen05.mma (N = 192 each) fill
// TMEM columns [0,192) and [192,384).  The accumulator cannot be double buffered (2 x 384 > 512 columns),
// so a tile's epilogue and the next tile's main loop run back to back; the TMA producer keeps
// prefetching the next tile's operand stages meanwhile.
// Epilogue (8 warps = two groups x 4 lane quarters; two threads per row, alternate 32-column chunks):
//   pass A  y = acc + bias + resid  -> written back to TMEM, partial row sums
//   pass B  partial sums of (y - mean)^2        (two-pass variance like nn.LayerNorm)
//   pass C  (y - mean) * rstd * gamma + beta -> four 256-bit global stores per thread and chunk
// Row statistics are exchanged between the two threads of a row through shared memory.
// H16 = true: fp16 operand planes (kind::f16 on the hi planes, 64 K-elements per 128-byte swizzle row; the accumulator is
// multiplied by the planes' exact inverse scale) for FS2_MATH_F16's out-projection and w_2; either variant can also emit
// the normalised rows as the hi plane (scaled by kPlaneScale) of the contraction that consumes them next.
#include "tc_common.cuh"

namespace fs2 {
namespace {
using namespace tc;

constexpr int BM = 128, BK = 32, UMMA_K = 8;
constexpr int A_BYTES = BM * BK * 4;
constexpr int LN_THREADS = 320;

template <int C, bool H16>
struct LCfg {
  static constexpr int HALF = C / 2;                      // one MMA / one TMA box of weight rows
  static constexpr int BKE = H16 ? 2 * BK : BK;            // K elements per pipeline step
  static constexpr int B_BYTES = C * BK * 4;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGING_BYTES = 2 * BM * 4;         // row-statistic exchange between the two threads of a row
  static constexpr int STAGES = (227 * 1024 - STAGING_BYTES - 1024 - 512) / STAGE_BYTES;
  static constexpr size_t SMEM = (size_t)STAGES * STAGE_BYTES + STAGING_BYTES + 1024 + 512;
  static constexpr uint32_t IDESC = H16 ? idesc_f16(BM, HALF) : idesc_tf32(BM, HALF);
  static constexpr int TMEM_COLS = 512;
  static constexpr int CHUNKS_PER_GROUP = C / 64;
  static_assert(C % 64 == 0 && HALF % 16 == 0 && HALF <= 256 && C <= 512, "row width");
  static_assert((HALF * BK * 4) % 1024 == 0 && STAGES >= 2, "smem layout");
};

struct LnParams {
  int M, K;
  const float* bias; const float* resid; int ldr;
  const float* gamma; const float* beta; float eps;
  float* out; int ldo;
  __half* out_h; int ldo_h;     // optional hi plane of the result (scaled by kPlaneScale)
  float a_inv; const float* w_inv;   // accumulator scale (operand planes' inverse pre-scale); 1 / null in the tf32 variant
};

template <int C, bool H16>
__global__ void __launch_bounds__(LN_THREADS, 1)
gemm_ln_tf32_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b, LnParams p) {
  using L = LCfg<C, H16>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* tiles = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // keeps the shared address space (no generic LD/ST)
  uint8_t* staging = tiles + (size_t)L::STAGES * L::STAGE_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(staging + L::STAGING_BYTES);
  uint64_t* empty_bar = full_bar + L::STAGES;
  uint64_t* acc_full = empty_bar + L::STAGES;
  uint64_t* acc_empty = acc_full + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 1);

  const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;
  const int steps = (p.K + L::BKE - 1) / L::BKE;
  const int tiles_total = (p.M + BM - 1) / BM;

  if (threadIdx.x == 0) {
    for (int s = 0; s < L::STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(acc_full, 1); mbar_init(acc_empty, 8);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc(tmem_slot, L::TMEM_COLS);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {  // ---- TMA producer ----
      int n = 0;
      for (int tile = blockIdx.x; tile < tiles_total; tile += gridDim.x) {
        const int r0 = tile * BM;
        for (int s = 0; s < steps; ++s, ++n) {
          const int slot = n % L::STAGES;
          mbar_wait(&empty_bar[slot], ((n / L::STAGES) & 1) ^ 1);
          uint8_t* st = tiles + (size_t)slot * L::STAGE_BYTES;
          mbar_expect_tx(&full_bar[slot], L::STAGE_BYTES);
          tma_load_3d(st, &tmap_a, &full_bar[slot], s * L::BKE, r0, 0);
          tma_load_3d(st + A_BYTES, &tmap_b, &full_bar[slot], s * L::BKE, 0, 0);
          tma_load_3d(st + A_BYTES + L::HALF * BK * 4, &tmap_b, &full_bar[slot], s * L::BKE, L::HALF, 0);
        }
      }
    }
  } else if (warp == 1) {
    // ---- MMA issuer: whole warp, one lane elected inside each tcgen05 asm ----
    int n = 0, it = 0;
    for (int tile = blockIdx.x; tile < tiles_total; tile += gridDim.x, ++it) {
      mbar_wait(acc_empty, (it & 1) ^ 1);            // the epilogue has drained the (single) accumulator
      tcgen05_fence_after();
      for (int s = 0; s < steps; ++s, ++n) {
        const int slot = n % L::STAGES;
        mbar_wait(&full_bar[slot], (n / L::STAGES) & 1);
        tcgen05_fence_after();
        const uint32_t base = smem_u32(tiles + (size_t)slot * L::STAGE_BYTES);
        const uint64_t a = make_sw128_kmajor_desc(base);
        const uint64_t b0 = make_sw128_kmajor_desc(base + A_BYTES), b1 = make_sw128_kmajor_desc(base + A_BYTES + L::HALF * BK * 4);
#pragma unroll
        for (int k = 0; k < BK / UMMA_K; ++k) {
          if (H16) {
            umma_f16(tmem_base, a + 2 * k, b0 + 2 * k, L::IDESC, (s | k) != 0);
            umma_f16(tmem_base + L::HALF, a + 2 * k, b1 + 2 * k, L::IDESC, (s | k) != 0);
          } else {
            umma_tf32(tmem_base, a + 2 * k, b0 + 2 * k, L::IDESC, (s | k) != 0);
            umma_tf32(tmem_base + L::HALF, a + 2 * k, b1 + 2 * k, L::IDESC, (s | k) != 0);
          }
        }
        tcgen05_commit(&empty_bar[slot]);
      }
      tcgen05_commit(acc_full);
    }
  } else {
    // ---- epilogue: group g = (warp-2)/4 takes the 32-column chunks g, g+2, g+4, ...; thread == row ----
    const int wq = warp & 3, grp = (warp - 2) >> 2;
    const int row = wq * 32 + lane;
    const uint32_t taddr = tmem_base + ((uint32_t)(wq * 32) << 16);
    float* xchg = reinterpret_cast<float*>(staging);          // [2][128] floats
    float v[32];
    int it = 0;
    const float oscale = p.a_inv * (p.w_inv ? __ldg(p.w_inv) : 1.0f);
    for (int tile = blockIdx.x; tile < tiles_total; tile += gridDim.x, ++it) {
      const int r0 = tile * BM;
      const long m = (long)r0 + row;
      const bool row_ok = m < p.M;
      // while the tensor core runs this tile's main loop, pull this thread's share of the residual row into L2
      // (6 x 128-byte lines): the pass-A loads then see L2 latency instead of an exposed HBM round trip per chunk
      if (p.resid && row_ok) {
#pragma unroll
        for (int i = 0; i < L::CHUNKS_PER_GROUP; ++i) prefetch_l2(p.resid + m * p.ldr + grp * 32 + i * 64);
      }
      mbar_wait(acc_full, it & 1);
      tcgen05_fence_after();
      // pass A: y = acc + bias + resid, kept in TMEM; partial row sum
      float sum = 0.f;
#pragma unroll 1
      for (int i = 0; i < L::CHUNKS_PER_GROUP; ++i) {
        const int c0 = grp * 32 + i * 64;
        float4 bv[8], rv[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          bv[q] = p.bias ? __ldg(reinterpret_cast<const float4*>(p.bias + c0 + q * 4)) : make_float4(0.f, 0.f, 0.f, 0.f);
          rv[q] = (p.resid && row_ok) ? __ldg(reinterpret_cast<const float4*>(p.resid + m * p.ldr + c0 + q * 4))
                                      : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        __syncwarp();
        tmem_ld32(taddr + c0, v);
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          v[q * 4 + 0] = fmaf(v[q * 4 + 0], oscale, bv[q].x + rv[q].x); v[q * 4 + 1] = fmaf(v[q * 4 + 1], oscale, bv[q].y + rv[q].y);
          v[q * 4 + 2] = fmaf(v[q * 4 + 2], oscale, bv[q].z + rv[q].z); v[q * 4 + 3] = fmaf(v[q * 4 + 3], oscale, bv[q].w + rv[q].w);
          sum += (v[q * 4] + v[q * 4 + 1]) + (v[q * 4 + 2] + v[q * 4 + 3]);
        }
        tmem_st32(taddr + c0, v);
      }
      tmem_st_wait();
      xchg[grp * BM + row] = sum;
      named_bar_sync(3, 256);
      const float mean = (sum + xchg[(grp ^ 1) * BM + row]) * (1.0f / C);
      named_bar_sync(3, 256);
      // pass B: partial sum of squared deviations
      float ssq = 0.f;
#pragma unroll 1
      for (int i = 0; i < L::CHUNKS_PER_GROUP; ++i) {
        __syncwarp();
        tmem_ld32(taddr + grp * 32 + i * 64, v);
#pragma unroll
        for (int j = 0; j < 32; ++j) { const float d = v[j] - mean; ssq = fmaf(d, d, ssq); }
      }
      xchg[grp * BM + row] = ssq;
      named_bar_sync(3, 256);
      const float rstd = 1.0f / sqrtf((ssq + xchg[(grp ^ 1) * BM + row]) * (1.0f / C) + p.eps);
      named_bar_sync(3, 256);                                  // xchg reads done before the next tile rewrites it
      // pass C: normalise, affine, stage, TMA store
#pragma unroll 1
      for (int i = 0; i < L::CHUNKS_PER_GROUP; ++i) {
        const int c0 = grp * 32 + i * 64;
        float4 gv[8], bt[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          gv[q] = __ldg(reinterpret_cast<const float4*>(p.gamma + c0 + q * 4));
          bt[q] = __ldg(reinterpret_cast<const float4*>(p.beta + c0 + q * 4));
        }
        __syncwarp();
        tmem_ld32(taddr + c0, v);
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          v[q * 4 + 0] = (v[q * 4 + 0] - mean) * rstd * gv[q].x + bt[q].x; v[q * 4 + 1] = (v[q * 4 + 1] - mean) * rstd * gv[q].y + bt[q].y;
          v[q * 4 + 2] = (v[q * 4 + 2] - mean) * rstd * gv[q].z + bt[q].z; v[q * 4 + 3] = (v[q * 4 + 3] - mean) * rstd * gv[q].w + bt[q].w;
        }
        if (row_ok) {
          float* dst = p.out + m * p.ldo + c0;
#pragma unroll
          for (int q = 0; q < 4; ++q) st_global_v8(dst + q * 8, v + q * 8);
          if (p.out_h != nullptr) {
            uint32_t h[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) h[j] = hi_pair(v[2 * j], v[2 * j + 1]);
            __half* dh = p.out_h + m * p.ldo_h + c0;
            st_global_v8_b32(dh, h); st_global_v8_b32(dh + 16, h + 8);
          }
        }
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(acc_empty);
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    tmem_dealloc(tmem_base, L::TMEM_COLS);
  }
}

}  // namespace

bool gemm_ln_tf32_supported(const TapGemm& g) { return g.taps == 1 && g.N == 384 && g.K % 4 == 0 && g.ln_gamma && !g.vt_out && g.act == ACT_NONE; }

namespace {
template <bool H16>
int launch_ln(const TapGemm& g, cudaStream_t st) {
  constexpr int C = 384;
  using L = LCfg<C, H16>;
  const uint64_t M = (uint64_t)g.B * g.L;
  static unsigned long long configured = 0;   // per-device bit mask
  int rc;
  if ((rc = ensure_smem_attr(gemm_ln_tf32_kernel<C, H16>, L::SMEM, &configured))) return rc;
  CUtensorMap ma, mb;
  const int esz = H16 ? 2 : 4;
  const uint64_t arow = H16 ? (uint64_t)g.K * 2 : (uint64_t)g.ldx * 4;
  if ((rc = make_map(&ma, H16 ? (const void*)g.xp : (const void*)g.x, g.K, M, 1, arow, arow * M, BM, H16))) return rc;
  if ((rc = make_map(&mb, H16 ? (const void*)g.w_hi : (const void*)g.w, g.K, C, 1, (uint64_t)g.K * esz, (uint64_t)g.K * esz * C, L::HALF, H16))) return rc;
  LnParams p;
  p.M = (int)M; p.K = g.K; p.bias = g.bias; p.resid = g.resid; p.ldr = g.ldr;
  p.gamma = g.ln_gamma; p.beta = g.ln_beta; p.eps = g.ln_eps; p.out = g.out; p.ldo = g.ldo;
  p.out_h = g.outp; p.ldo_h = g.ldo_p;
  p.a_inv = H16 ? g.a_inv : 1.0f; p.w_inv = H16 ? g.w_inv : nullptr;
  const int tiles = (int)((M + BM - 1) / BM);
  const int grid = tiles < sm_count_current() ? tiles : sm_count_current();
  gemm_ln_tf32_kernel<C, H16><<<grid, LN_THREADS, L::SMEM, st>>>(ma, mb, p);
  FS2_LAUNCH_CHECK();
  return FS2_OK;
}
}  // namespace

// g.xp set: fp16 operand planes (hi planes of x and w); g.outp set: also emit the result's hi plane
int gemm_ln_tf32(const TapGemm& g, cudaStream_t st) {
  FS2_REQUIRE(gemm_ln_tf32_supported(g), "gemm_ln_tf32: unsupported shape (N=%d taps=%d)", g.N, g.taps);
  FS2_REQUIRE(g.ldo % 8 == 0 && (!g.resid || g.ldr % 4 == 0) && g.out && (reinterpret_cast<uintptr_t>(g.out) & 31) == 0,
              "gemm_ln_tf32: row strides / output alignment");
  FS2_REQUIRE(!g.outp || (g.ldo_p % 16 == 0 && (reinterpret_cast<uintptr_t>(g.outp) & 31) == 0), "gemm_ln_tf32: fp16 output rows must be 32-byte aligned");
  if ((uint64_t)g.B * g.L == 0) return FS2_OK;
  if (g.xp) {
    FS2_REQUIRE(g.w_hi && g.K % 8 == 0 && (reinterpret_cast<uintptr_t>(g.xp) & 15) == 0 && (reinterpret_cast<uintptr_t>(g.w_hi) & 15) == 0,
                "gemm_ln_tf32: fp16 operands must be 16-byte aligned with K a multiple of 8");
    return launch_ln<true>(g, st);
  }
  FS2_REQUIRE(g.ldx % 4 == 0, "gemm_ln_tf32: row strides / output alignment");
  return launch_ln<false>(g, st);
}

}  // namespace fs2
