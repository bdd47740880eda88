// Tensor-core multi-head self-attention core for sm_100a (core/attention.py:52-73):
// S = Q K^T and O = P V as tcgen05.mma kind::tf32 with S, P and O resident in TMEM, Q/K/V^T tiles
// staged by TMA.  The [B,h,L,L] score tensor never touches HBM.
//
//   ctx[b,t,h*dk:(h+1)*dk] = softmax_u( q.k_u / sqrt(dk) | u < len_b ) . v ,  0 for t >= len_b
//   (lens == nullptr: no masking at all -- the reference's mask=None branch, attention.py:67)
//
// Operands (all fp32, consumed as tf32 by the tensor core):
//   Q, K : rows of the fused projection buffer qkv [B, L, 3C]  (K-major: dk contiguous)
//   V^T  : vt [B*heads, dk, lpad], written transposed by the projection GEMM's epilogue so that the
//          P.V product has a K-major B operand too (kv contiguous)
//   P    : written by the softmax warps straight into TMEM and consumed from there as the A operand.
//
// CTA = 128 queries of one (batch, head).  Softmax is single-pass ("online") with a *lazy* reference
// maximum: P = exp2(c*(s - m_ref)) is written into TMEM in place of S and O += P V^T accumulates in TMEM;
// m_ref only moves (and O, l are rescaled by exp2(c*(m_ref_old - m_ref_new)) through tcgen05.ld/st) when a
// tile's row maximum exceeds it by more than 2^8 in the exp2 domain, which after the first tile is rare, so the
// accumulator is almost never touched.  O / l at the end is exact for any reference (no overflow: P <= 2^8).
//
// Warp roles (10 warps): 0 = TMA producer, 1 = TMEM allocator + MMA issuer, 2..9 = softmax /
// epilogue (two threads per query row == TMEM lane, 64 score columns each; row statistics are
// combined through shared memory once per pass).  TMEM: two S/P buffers (2 x 128 columns) so the
// tensor core computes S_{j+1} while the softmax warps work on tile j, plus DK columns of O.
// Shared memory: Q resident (DK/32 boxes of 16 KB) + a ring of 24 KB slots for K / V^T boxes.
#include <math.h>

#include "tc_common.cuh"

namespace fs2 {
namespace {
using namespace tc;

constexpr int BQ = 128, BKV = 128, CH = 32;   // CH: fp32 per 128-byte swizzle row
constexpr int ATT_THREADS = 320;   // TMA, MMA, 8 softmax / epilogue warps

template <int DK>
struct ACfg {
  static constexpr int QCH = DK / CH;                  // K-chunks of the S product
  static constexpr int Q_BYTES = QCH * BQ * CH * 4;    // resident Q
  static constexpr int K_BOX = BKV * CH * 4;           // 16 KB: 128 kv rows x 32 dk
  static constexpr int V_BOX = DK * CH * 4;            // DK rows x 32 kv
  static constexpr int SLOT = V_BOX > K_BOX ? V_BOX : K_BOX;
  static constexpr int SLOTS = (211 * 1024 - Q_BYTES) / SLOT;
  static constexpr size_t SMEM = (size_t)Q_BYTES + (size_t)SLOTS * SLOT + 1024 + 512 + 4 * BQ * 4;
  static constexpr uint32_t IDESC_S = idesc_tf32(BQ, BKV);
  static constexpr uint32_t IDESC_O = idesc_tf32(BQ, DK);
  static constexpr int TMEM_COLS = 512;
  static constexpr int O_COL = 2 * BKV;                // S/P buffers at columns 0 and 128
  static_assert(DK % 64 == 0 && DK <= 256, "d_k");
  static_assert(SLOT % 1024 == 0 && SLOTS >= 3, "ring");
};

struct AParams {
  const int64_t* lens; int L, C, heads; float* ctx; float scale_log2e;
};

template <int DK>
__global__ void __launch_bounds__(ATT_THREADS, 1)
attention_tf32_kernel(const __grid_constant__ CUtensorMap tmap_qk, const __grid_constant__ CUtensorMap tmap_vt, AParams p) {
  using A = ACfg<DK>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* q_smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // keeps the shared address space
  uint8_t* ring = q_smem + A::Q_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(ring + (size_t)A::SLOTS * A::SLOT);
  uint64_t* full_bar = bars;                   // [SLOTS]
  uint64_t* empty_bar = bars + A::SLOTS;       // [SLOTS]
  uint64_t* q_bar = bars + 2 * A::SLOTS;       // Q landed
  uint64_t* s_full = q_bar + 1;                // [2] MMA -> softmax: S tile ready
  uint64_t* pv_done = s_full + 2;              // [2] (only [0] used) MMA -> softmax: P.V of the previous tile has finished
  uint64_t* p_full = pv_done + 2;              // [2] softmax -> MMA: P written
  uint64_t* o_full = p_full + 2;               // all P.V done
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_full + 1);
  float* xchg = reinterpret_cast<float*>(tmem_slot + 4);   // [2 tiles][2 halves][BQ] row-statistic exchange between column halves

  const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * BQ, h = blockIdx.y, b = blockIdx.z;
  const int len = p.lens ? (int)min((long)p.lens[b], (long)p.L) : p.L;   // keys >= len are masked
  const int J = (len + BKV - 1) / BKV;                                   // kv tiles that contain valid keys

  if (threadIdx.x == 0) {
    for (int s = 0; s < A::SLOTS; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(q_bar, 1);
    for (int i = 0; i < 2; ++i) { mbar_init(&s_full[i], 1); mbar_init(&pv_done[i], 1); mbar_init(&p_full[i], 8); }
    mbar_init(o_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc(tmem_slot, A::TMEM_COLS);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (J > 0) {
    if (warp == 0) {
      if (lane == 0) {  // ---- TMA producer: Q once, then K / V^T boxes in exactly the order the MMA warp consumes them ----
        mbar_expect_tx(q_bar, A::Q_BYTES);
        for (int c = 0; c < A::QCH; ++c) tma_load_3d(q_smem + (size_t)c * BQ * CH * 4, &tmap_qk, q_bar, h * DK + c * CH, q0, b);
        int n = 0;
        auto push_k = [&](int j) {
          for (int c = 0; c < A::QCH; ++c, ++n) {
            const int slot = n % A::SLOTS;
            mbar_wait(&empty_bar[slot], ((n / A::SLOTS) & 1) ^ 1);
            mbar_expect_tx(&full_bar[slot], A::K_BOX);
            tma_load_3d(ring + (size_t)slot * A::SLOT, &tmap_qk, &full_bar[slot], p.C + h * DK + c * CH, j * BKV, b);
          }
        };
        auto push_v = [&](int j) {
          for (int c = 0; c < BKV / CH; ++c, ++n) {
            const int slot = n % A::SLOTS;
            mbar_wait(&empty_bar[slot], ((n / A::SLOTS) & 1) ^ 1);
            mbar_expect_tx(&full_bar[slot], A::V_BOX);
            tma_load_3d(ring + (size_t)slot * A::SLOT, &tmap_vt, &full_bar[slot], j * BKV + c * CH, 0, b * p.heads + h);
          }
        };
        push_k(0);
        for (int j = 0; j < J; ++j) { if (j + 1 < J) push_k(j + 1); push_v(j); }
      }
    } else if (warp == 1) {
      {  // ---- MMA issuer: all 32 lanes run the loop, one lane is elected inside each tcgen05 asm ----
        mbar_wait(q_bar, 0);
        tcgen05_fence_after();
        const uint32_t q_addr = smem_u32(q_smem);
        int n = 0;
        // S tile j goes to S/P buffer j & 1; S_{j+2} reuses it after P.V_j, which is issued earlier in this thread
        auto issue_s = [&](int g) {
          const uint32_t d = tmem_base + (uint32_t)((g & 1) * BKV);
          for (int c = 0; c < A::QCH; ++c, ++n) {
            const int slot = n % A::SLOTS;
            mbar_wait(&full_bar[slot], (n / A::SLOTS) & 1);
            tcgen05_fence_after();
            const uint64_t adesc = make_sw128_kmajor_desc(q_addr + c * BQ * CH * 4);
            const uint64_t bdesc = make_sw128_kmajor_desc(smem_u32(ring + (size_t)slot * A::SLOT));
#pragma unroll
            for (int k = 0; k < CH / 8; ++k) umma_tf32(d, adesc + 2 * k, bdesc + 2 * k, A::IDESC_S, (c | k) != 0);
            tcgen05_commit(&empty_bar[slot]);
          }
          tcgen05_commit(&s_full[g & 1]);
        };
        issue_s(0);
        for (int j = 0; j < J; ++j) {
          if (j + 1 < J) issue_s(j + 1);                 // the tensor core computes S_{j+1} while the softmax warps work on tile j
          mbar_wait(&p_full[j & 1], (j >> 1) & 1);
          tcgen05_fence_after();
          const uint32_t p_tmem = tmem_base + (uint32_t)((j & 1) * BKV);
          for (int c = 0; c < BKV / CH; ++c, ++n) {
            const int slot = n % A::SLOTS;
            mbar_wait(&full_bar[slot], (n / A::SLOTS) & 1);
            tcgen05_fence_after();
            const uint64_t bdesc = make_sw128_kmajor_desc(smem_u32(ring + (size_t)slot * A::SLOT));
#pragma unroll
            for (int k = 0; k < CH / 8; ++k)
              umma_tf32_ts(tmem_base + A::O_COL, p_tmem + c * CH + k * 8, bdesc + 2 * k, A::IDESC_O, (j | c | k) != 0);
            tcgen05_commit(&empty_bar[slot]);
          }
          tcgen05_commit(&pv_done[0]);                   // lets the softmax warps rescale O if tile j+1 raises the reference max
        }
        tcgen05_commit(o_full);
      }
    } else {
      // ---- softmax / epilogue: 8 warps; warps w and w+4 share TMEM lane quarter w%4 and split the columns ----
      const int wq = warp & 3, half = (warp - 2) >> 2;          // half 0: columns [0,64), half 1: [64,128)
      const int row = wq * 32 + lane;
      const uint32_t lane_addr = tmem_base + ((uint32_t)(wq * 32) << 16);
      float v[64];
      float m_ref = -INFINITY, l_row = 0.f;     // reference maximum (raw-score domain) and row sum relative to it
      const float c_exp = p.scale_log2e;
      for (int j = 0; j < J; ++j) {
        mbar_wait(&s_full[j & 1], (j >> 1) & 1);
        tcgen05_fence_after();
        const int kv0 = j * BKV + half * 64;
        __syncwarp();
        const uint32_t ta = lane_addr + (uint32_t)((j & 1) * BKV + half * 64);
        tmem_ld32_nowait(ta, v); tmem_ld32_nowait(ta + 32, v + 32); tmem_ld_wait_pin<64>(v);
        float tmax = -INFINITY;
#pragma unroll
        for (int i = 0; i < 64; ++i) if (kv0 + i < len) tmax = fmaxf(tmax, v[i]);
        float* xr = xchg + (j & 1) * 2 * BQ;                     // double-buffered exchange: one barrier per tile
        xr[half * BQ + row] = tmax;
        named_bar_sync(1, 256);
        tmax = fmaxf(tmax, xr[(half ^ 1) * BQ + row]);           // both threads of the row now hold the tile's row maximum
        // lazy reference update: move m_ref only if the tile exceeds it by more than 2^8 in the exp2 domain
        const bool bump = (tmax - m_ref) * c_exp > 8.0f;          // first tile: m_ref = -inf -> always
        if (__any_sync(0xffffffffu, bump)) {
          const float alpha = bump ? fast_exp2((m_ref - tmax) * c_exp) : 1.0f;   // exp2(-inf) = 0 on the first tile
          if (j > 0) {   // O holds tiles 0..j-1: wait until P.V_{j-1} has landed, then scale this thread's half of the row
            mbar_wait(&pv_done[0], (j - 1) & 1);
            tcgen05_fence_after();
            float o[32];
#pragma unroll 1
            for (int c0 = 0; c0 < DK / 2; c0 += 32) {
              const uint32_t oa = lane_addr + (uint32_t)(A::O_COL + half * (DK / 2) + c0);
              __syncwarp();
              tmem_ld32(oa, o);
#pragma unroll
              for (int i = 0; i < 32; ++i) o[i] *= alpha;
              tmem_st32(oa, o);
            }
            tmem_st_wait();
          }
          l_row *= alpha;
          if (bump) m_ref = tmax;
        }
        const float mb = m_ref * c_exp;
#pragma unroll
        for (int i = 0; i < 64; ++i) {
          const float e = (kv0 + i < len) ? fast_exp2(fmaf(v[i], c_exp, -mb)) : 0.f;
          v[i] = e; l_row += e;
        }
        tmem_st32(ta, v); tmem_st32(ta + 32, v + 32);
        tmem_st_wait();
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_full[j & 1]);
      }
      named_bar_sync(1, 256);                                     // last exchange buffer is free again
      xchg[half * BQ + row] = l_row;
      named_bar_sync(1, 256);
      l_row += xchg[(half ^ 1) * BQ + row];
      // epilogue: O / l -> ctx rows (0 for masked query rows); each half stores DK/2 columns
      mbar_wait(o_full, 0);
      tcgen05_fence_after();
      const int t = q0 + row;
      const bool store = t < p.L;
      const float inv = (p.lens && t >= len) ? 0.f : 1.0f / l_row;
      float* dst = p.ctx + ((long)b * p.L + t) * p.C + h * DK + half * (DK / 2);
#pragma unroll 1
      for (int c0 = 0; c0 < DK / 2; c0 += 32) {
        __syncwarp();
        tmem_ld32(lane_addr + (uint32_t)(A::O_COL + half * (DK / 2) + c0), v);
        if (store) {
#pragma unroll
          for (int q = 0; q < 8; ++q)
            *reinterpret_cast<float4*>(dst + c0 + q * 4) = make_float4(v[q * 4] * inv, v[q * 4 + 1] * inv, v[q * 4 + 2] * inv, v[q * 4 + 3] * inv);
        }
      }
    }
  } else if (warp >= 2) {
    // no valid key at all (len == 0): the reference's masked_fill turns the NaN rows into 0
    const int t = q0 + (warp & 3) * 32 + lane;
    if (t < p.L && warp < 6) {
      float* dst = p.ctx + ((long)b * p.L + t) * p.C + h * DK;
      for (int c = 0; c < DK; c += 4) *reinterpret_cast<float4*>(dst + c) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    tmem_dealloc(tmem_base, A::TMEM_COLS);
  }
}

template <int DK>
int launch(const float* qkv, const float* vt, int lpad, const int64_t* lens, int B, int L, int C, int heads, float* ctx,
           cudaStream_t st) {
  using A = ACfg<DK>;
  static bool configured = false;
  if (!configured) {
    FS2_CUDA_CHECK(cudaFuncSetAttribute(attention_tf32_kernel<DK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)A::SMEM));
    configured = true;
  }
  CUtensorMap mqk, mvt;
  int rc;
  const uint64_t row = (uint64_t)3 * C * 4;
  if ((rc = make_map(&mqk, qkv, (uint64_t)3 * C, L, B, row, row * L, BQ))) return rc;
  // extent L (not lpad) along kv: alignment padding columns are never read, TMA zero-fills past L
  if ((rc = make_map(&mvt, vt, L, DK, (uint64_t)B * heads, (uint64_t)lpad * 4, (uint64_t)lpad * 4 * DK, DK))) return rc;
  AParams p;
  p.lens = lens; p.L = L; p.C = C; p.heads = heads; p.ctx = ctx;
  p.scale_log2e = (1.0f / sqrtf((float)DK)) * 1.4426950408889634f;
  dim3 grid((L + BQ - 1) / BQ, heads, B);
  attention_tf32_kernel<DK><<<grid, ATT_THREADS, A::SMEM, st>>>(mqk, mvt, p);
  FS2_LAUNCH_CHECK();
  return FS2_OK;
}

__global__ void transpose_v_kernel(const float* __restrict__ qkv, int L, int C, float* __restrict__ vt, int lpad) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z, t0 = blockIdx.x * 32, n0 = blockIdx.y * 32;   // n over the C columns of the V third
  int t = t0 + threadIdx.y, n = n0 + threadIdx.x;
  tile[threadIdx.y][threadIdx.x] = (t < L) ? qkv[((long)b * L + t) * 3 * C + 2 * C + n] : 0.f;
  __syncthreads();
  t = t0 + threadIdx.x; n = n0 + threadIdx.y;
  if (t < L) vt[((long)b * C + n) * lpad + t] = tile[threadIdx.x][threadIdx.y];   // (b*heads + h)*dk + d == b*C + n
}

}  // namespace

int transpose_v(const float* qkv, int B, int L, int C, int heads, float* vt, int lpad, cudaStream_t st) {
  (void)heads;
  dim3 grid((L + 31) / 32, C / 32, B), block(32, 32);
  transpose_v_kernel<<<grid, block, 0, st>>>(qkv, L, C, vt, lpad);
  FS2_LAUNCH_CHECK();
  return FS2_OK;
}

int attention_tf32(const float* qkv, const float* vt, int lpad, const int64_t* lens, int B, int L, int C, int heads,
                   float* ctx, cudaStream_t st) {
  FS2_REQUIRE(heads > 0 && C % heads == 0, "attention: C=%d not divisible by heads=%d", C, heads);
  FS2_REQUIRE(vt && lpad >= L && lpad % 4 == 0, "attention_tf32: needs the transposed V buffer with a 16-byte aligned row pitch");
  if (B == 0 || L == 0) return FS2_OK;
  const int dk = C / heads;
  if (dk == 192) return launch<192>(qkv, vt, lpad, lens, B, L, C, heads, ctx, st);
  if (dk == 128) return launch<128>(qkv, vt, lpad, lens, B, L, C, heads, ctx, st);
  set_error("attention_tf32: d_k=%d unsupported (128 or 192)", dk);
  return FS2_ERR_INVALID;
}

}  // namespace fs2
