// Tensor-core multi-head self-attention core on fp16 operand planes (core/attention.py:52-73), sm_100a:
//   F16 : S = Q K^T and O = P V as tcgen05.mma kind::f16 on the hi planes (FS2_MATH_F16's decoder);
//   X3  : the error-compensated form -- Q, K, V^T and P each as fp16 hi + lo, three products per term
//         (lo.hi + hi.lo + hi.hi), fp32 accumulation in TMEM, fp32 softmax statistics: fp32-class scores and context
//         (the encoder in every tensor-core mode, the decoder in FS2_MATH_3XTF32).
// S, P and O live in TMEM; the [B,h,L,L] score tensor never touches HBM.
//
//   ctx[b,t,h*dk:(h+1)*dk] = softmax_u( q.k_u / sqrt(dk) | u < len_b ) . v ,  0 for t >= len_b
//   (lens == nullptr: no masking at all -- the reference's mask=None branch, attention.py:67)
//
// Operands (written by the q|k|v projection's epilogue, gemm_tc.cu; all scaled by kPlaneScale):
//   Q, K : planes qkp [P][B*L][2C] fp16 (q | k, heads contiguous inside each; K-major: dk contiguous)
//   V^T  : planes vtp [P][B*heads][dk][lpad] fp16, i.e. V stored transposed so that P.V has a K-major B operand too
//   P    : written by the softmax warps straight into TMEM as packed fp16 pairs -- in place of the fp32 S tile: hi in
//          columns [0,64), lo in [64,128) of the tile's 128 columns -- and consumed from there as the A operand.
// Result: the context as the operand planes of the out-projection (hi, + lo in X3) and / or as fp32 rows.
//
// CTA = 128 queries of one (batch, head); 128 keys per step.  Softmax is single-pass ("online") with a *lazy* reference
// maximum: m_ref only moves (and O, l are rescaled through tcgen05.ld/st) when a tile's row maximum exceeds it by more
// than 2^8 in the exp2 domain, which after the first tile is rare.  O / l at the end is exact for any reference
// (no overflow: P <= 2^8, well inside fp16).  In F16 the row sum l adds the *rounded* P values, so the weights the
// tensor core applies still sum to one.
//
// Warp roles (10 warps): 0 = TMA producer, 1 = TMEM allocator + MMA issuer, 2..9 = softmax / epilogue (two threads per
// query row == TMEM lane, 64 score columns each).  TMEM: two S/P buffers (2 x 128 columns) so the tensor core computes
// S_{j+1} while the softmax warps work on tile j, plus DK columns of O.  Shared memory: Q resident (P x DK/64 boxes of
// 16 KB) + a ring of K / V^T boxes (one 128-byte swizzle row = 64 fp16 wide).
#include <math.h>
#include <stdlib.h>

#include "tc_common.cuh"

namespace fs2 {
namespace {
using namespace tc;

constexpr int BQ = 128, BKV = 128, CH = 64;   // CH: fp16 per 128-byte swizzle row
constexpr int ATT_THREADS = 320;              // TMA, MMA, 8 softmax / epilogue warps

template <int DK, bool X3>
struct HCfg {
  static constexpr int P = X3 ? 2 : 1;                 // operand planes
  static constexpr int QCH = DK / CH;                  // K-chunks of the S product
  static constexpr int Q_BOX = BQ * 128;               // 16 KB: 128 query rows x 64 dk
  static constexpr int Q_BYTES = P * QCH * Q_BOX;      // resident Q
  static constexpr int K_BOX = BKV * 128;              // 16 KB: 128 kv rows x 64 dk
  static constexpr int V_BOX = DK * 128;               // DK rows x 64 kv
  static constexpr int SLOT = V_BOX > K_BOX ? V_BOX : K_BOX;
  static constexpr int SLOTS_MAX = (222 * 1024 - Q_BYTES) / SLOT;
  static constexpr int SLOTS = SLOTS_MAX > 8 ? 8 : SLOTS_MAX;
  static constexpr size_t SMEM = (size_t)Q_BYTES + (size_t)SLOTS * SLOT + 1024 + 512 + 4 * BQ * 4;
  static constexpr uint32_t IDESC_S = idesc_f16(BQ, BKV);
  static constexpr uint32_t IDESC_O = idesc_f16(BQ, DK);
  static constexpr int TMEM_COLS = 512;
  static constexpr int O_COL = 2 * BKV;                // S/P buffers at columns 0 and 128
  static_assert(DK % 64 == 0 && DK <= 256, "d_k");
  static_assert(SLOT % 1024 == 0 && SLOTS >= 3, "ring");
};

struct HParams {
  const int64_t* lens; int B, L, C, heads;
  float* ctx;            // [B,L,C] fp32 or null
  __half* ctxp;          // planes [P][B*L][C] or null
  float scale_log2e;     // log2(e) / sqrt(dk) / kPlaneScale^2  (Q and K are both pre-scaled)
  int debug;             // FS2_ATT_DEBUG (profiling experiments only): 1 skip exp math, 2 skip the S load, 4 skip the P store
  long long* trace;      // FS2_ATT_TRACE=1 (profiling experiments only): clock64 stamps of CTA (0,0,0), see att_trace_dump
};
constexpr int TRACE_SLOTS = 8, TRACE_TILES = 32;     // per tile: 2 softmax halves x 8 stamps, then 8 MMA-warp stamps
__device__ __forceinline__ void trace_at(const HParams& p, int j, int who, int k) {
  if (p.trace != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && (threadIdx.x & 31) == 0 && j < TRACE_TILES)
    p.trace[(j * 3 + who) * TRACE_SLOTS + k] = clock64();
}
// FS2_ATT_TRACE=1: a device buffer for the clock64 stamps of CTA (0,0,0); dumped (with a stream sync) after the launch
inline long long* att_trace_buffer() {
  static int on = -1; static long long* buf = nullptr;
  if (on < 0) { const char* e = getenv("FS2_ATT_TRACE"); on = e ? atoi(e) : 0; }
  if (on && !buf) { cudaMalloc(&buf, 32 * 3 * 8 * sizeof(long long)); }
  if (on && buf) cudaMemset(buf, 0, 32 * 3 * 8 * sizeof(long long));
  return on ? buf : nullptr;
}
inline void att_trace_dump(long long* dev, int tiles, cudaStream_t st) {
  static int dumped = 0;
  if (dumped++ >= 2) return;
  long long h[32 * 3 * 8];
  cudaStreamSynchronize(st);
  cudaMemcpy(h, dev, sizeof(h), cudaMemcpyDeviceToHost);
  const long long t0 = h[(0 * 3 + 2) * 8 + 0] ? h[(0 * 3 + 2) * 8 + 0] : h[0];
  fprintf(stderr, "fs2 attention trace (cycles since the MMA warp's first stamp; CTA 0)\n tile | softmax half0: wait_begin s_ready ld_done max_done exp_done p_arrived | half1: ... | mma: loop_top s_next_issued p_ready pv_issued\n");
  for (int j = 0; j < tiles && j < 32; ++j) {
    fprintf(stderr, " %3d |", j);
    for (int who = 0; who < 3; ++who) {
      for (int k = 0; k < (who == 2 ? 4 : 6); ++k) fprintf(stderr, " %7lld", h[(j * 3 + who) * 8 + k] ? h[(j * 3 + who) * 8 + k] - t0 : -1);
      fprintf(stderr, " |");
    }
    fprintf(stderr, "\n");
  }
}
inline int att_debug() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("FS2_ATT_DEBUG"); v = e ? atoi(e) : 0; }
  return v;
}

__device__ __forceinline__ void tmem_st32u(uint32_t taddr, const uint32_t* r) { tmem_st32(taddr, reinterpret_cast<const float*>(r)); }

// exp2 of one thread's 2 NP scores (64, or 32 in a short last tile) -> packed fp16 P (hi, and lo when X3); returns the
// row-sum contribution
template <bool X3, bool MASKED, int NP = 32>
__device__ __forceinline__ float softmax_tile(const float* v, float c_exp, float mb, int kv0, int len, uint32_t* ph, uint32_t* pl) {
  float s0 = 0.f, s1 = 0.f;
#pragma unroll
  for (int i = 0; i < NP; ++i) {
    float e0 = fast_exp2(fmaf(v[2 * i], c_exp, -mb)), e1 = fast_exp2(fmaf(v[2 * i + 1], c_exp, -mb));
    if (MASKED) { if (kv0 + 2 * i >= len) e0 = 0.f; if (kv0 + 2 * i + 1 >= len) e1 = 0.f; }
    const __half2 h = __floats2half2_rn(e0, e1);
    ph[i] = *reinterpret_cast<const uint32_t*>(&h);
    if (X3) {
      const float2 g = __half22float2(h);
      const __half2 l = __floats2half2_rn(e0 - g.x, e1 - g.y);
      pl[i] = *reinterpret_cast<const uint32_t*>(&l);
      s0 += e0; s1 += e1;
    } else {               // the tensor core sees the rounded weights: normalise by their sum
      const float2 g = __half22float2(h);
      s0 += g.x; s1 += g.y;
    }
  }
  return s0 + s1;
}

template <int DK, bool X3>
__global__ void __launch_bounds__(ATT_THREADS, 1)
attention_f16_kernel(const __grid_constant__ CUtensorMap tmap_qk, const __grid_constant__ CUtensorMap tmap_vt, HParams p) {
  using A = HCfg<DK, X3>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* q_smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // keeps the shared address space
  uint8_t* ring = q_smem + A::Q_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(ring + (size_t)A::SLOTS * A::SLOT);
  uint64_t* full_bar = bars;                   // [SLOTS]
  uint64_t* empty_bar = bars + A::SLOTS;       // [SLOTS]
  uint64_t* q_bar = bars + 2 * A::SLOTS;       // TMA -> MMA: Q of the current work item landed
  uint64_t* s_full = q_bar + 1;                // [2] MMA -> softmax: S tile ready
  uint64_t* pv_done = s_full + 2;              // MMA -> softmax: P.V of the previous tile has finished
  uint64_t* p_full = pv_done + 2;              // [2] softmax -> MMA: P written
  uint64_t* o_full = p_full + 2;               // MMA -> epilogue: all P.V of the work item done
  uint64_t* o_free = o_full + 1;               // epilogue -> MMA: O has been read, the next work item may overwrite it
  uint64_t* q_free = o_free + 1;               // MMA -> TMA: every S product of the work item has read Q
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(q_free + 1);
  float* xchg = reinterpret_cast<float*>(tmem_slot + 4);   // [2 tiles][2 halves][BQ] row-statistic exchange between column halves

  const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;
  // Persistent CTA: work item w = ((b * heads + h) * nq + qt), taken round-robin.  The TMA producer and the MMA warp run
  // ahead into the next item (its Q lands while the current item's last softmax / P.V / epilogue are still running), so
  // the 2-3 us cold-start of a tile (Q + first K over TMA, first S) is paid once per CTA instead of once per tile.
  const int nq = (p.L + BQ - 1) / BQ;
  const int total_work = p.B * p.heads * nq;
  auto decode = [&](int w, int& b, int& h, int& q0, int& len, int& J) {
    const int qt = w % nq; const int bh = w / nq;
    h = bh % p.heads; b = bh / p.heads; q0 = qt * BQ;
    len = p.lens ? (int)min((long)p.lens[b], (long)p.L) : p.L;       // keys >= len are masked
    J = (len + BKV - 1) / BKV;                                        // kv tiles that contain valid keys
  };
  // keys of kv tile j the tensor core has to look at: all 128, or -- in the last tile of an utterance -- the valid ones
  // rounded up to the MMA's granularity of 16 (the softmax writes P = 0 for the masked keys inside that granule).  The S
  // product of a partial tile uses N = n16, its P.V only the first n16 / 16 K-steps and ceil(n16 / 64) V^T boxes.
  auto keys16 = [&](int j, int len) { const int r = len - j * BKV; return r >= BKV ? BKV : (r + 15) & ~15; };

  if (threadIdx.x == 0) {
    for (int s = 0; s < A::SLOTS; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(q_bar, 1);
    for (int i = 0; i < 2; ++i) { mbar_init(&s_full[i], 1); mbar_init(&pv_done[i], 1); mbar_init(&p_full[i], 8); }
    mbar_init(o_full, 1); mbar_init(o_free, 8); mbar_init(q_free, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc(tmem_slot, A::TMEM_COLS);
  pdl_wait();                                  // no global memory is touched above: it overlaps the previous kernel's tail
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    {  // ---- TMA producer: per work item Q once, then K / V^T boxes in exactly the order the MMA warp consumes them.  The whole
       // warp runs the loop, one lane is elected inside each asm, and a box's coordinates and addresses are computed before
       // the wait for its slot (gemm_tc.cu: what sits between "slot free" and "load issued" is refill latency) ----
      const uint32_t ring_addr = smem_u32(ring), full_addr = smem_u32(full_bar), q_addr0 = smem_u32(q_smem), qbar_addr = smem_u32(q_bar);
      int n = 0, wc = 0;    // ring position and count of non-empty work items, both run across items
      for (int w = blockIdx.x; w < total_work; w += gridDim.x) {
        int b, h, q0, len, J;
        decode(w, b, h, q0, len, J);
        if (J == 0) continue;
        if (wc > 0) mbar_wait(q_free, (wc - 1) & 1);      // the previous item's S products are done with Q
        mbar_expect_tx_elect(qbar_addr, A::Q_BYTES);
        for (int pl = 0; pl < A::P; ++pl)
          for (int c = 0; c < A::QCH; ++c)
            tma_load_3d_elect(q_addr0 + (uint32_t)((pl * A::QCH + c) * A::Q_BOX), &tmap_qk, qbar_addr, h * DK + c * CH, q0, b + pl * p.B);
        auto push_k = [&](int j) {
          for (int c = 0; c < A::QCH; ++c)
            for (int pl = 0; pl < A::P; ++pl, ++n) {
              const int slot = n % A::SLOTS;
              const uint32_t dst = ring_addr + (uint32_t)slot * A::SLOT, fb = full_addr + (uint32_t)slot * 8u;
              const int c0 = p.C + h * DK + c * CH, c2 = b + pl * p.B;
              pin_before(dst, fb, c0, c2);
              mbar_wait(&empty_bar[slot], ((n / A::SLOTS) & 1) ^ 1);
              mbar_expect_tx_elect(fb, A::K_BOX);
              tma_load_3d_elect(dst, &tmap_qk, fb, c0, j * BKV, c2);
            }
        };
        auto push_v = [&](int j) {
          const int vch = (keys16(j, len) + CH - 1) / CH;
          for (int c = 0; c < vch; ++c)
            for (int pl = 0; pl < A::P; ++pl, ++n) {
              const int slot = n % A::SLOTS;
              const uint32_t dst = ring_addr + (uint32_t)slot * A::SLOT, fb = full_addr + (uint32_t)slot * 8u;
              const int c0 = j * BKV + c * CH, c2 = b * p.heads + h + pl * p.B * p.heads;
              pin_before(dst, fb, c0, c2);
              mbar_wait(&empty_bar[slot], ((n / A::SLOTS) & 1) ^ 1);
              mbar_expect_tx_elect(fb, A::V_BOX);
              tma_load_3d_elect(dst, &tmap_vt, fb, c0, 0, c2);
            }
        };
        push_k(0);
        for (int j = 0; j < J; ++j) { if (j + 1 < J) push_k(j + 1); push_v(j); }
        ++wc;
      }
    }
  } else if (warp == 1) {
    {  // ---- MMA issuer: all 32 lanes run the loop, one lane is elected inside each tcgen05 asm ----
      const uint32_t q_addr = smem_u32(q_smem);
      int n = 0, wc = 0, g0 = 0;      // ring position, non-empty work items, kv tiles issued so far (all run across items)
      auto take = [&]() -> uint64_t {   // next ring slot, as a descriptor
        const int slot = n % A::SLOTS;
        mbar_wait(&full_bar[slot], (n / A::SLOTS) & 1);
        tcgen05_fence_after();
        return make_sw128_kmajor_desc(smem_u32(ring + (size_t)slot * A::SLOT));
      };
      auto release = [&]() { tcgen05_commit(&empty_bar[n % A::SLOTS]); ++n; };
      // S tile g goes to S/P buffer g & 1; S_{g+2} reuses it after P.V_g, which is issued earlier in this thread
      auto issue_s = [&](int g, int n16) {
        const uint32_t d = tmem_base + (uint32_t)((g & 1) * BKV);
        const uint32_t idesc_s = (A::IDESC_S & ~(0x3Fu << 17)) | ((uint32_t)(n16 >> 3) << 17);   // N = n16
        for (int c = 0; c < A::QCH; ++c) {
          const uint64_t q_hi = make_sw128_kmajor_desc(q_addr + c * A::Q_BOX);
          const uint64_t k_hi = take();
          if (X3) {
            const uint64_t q_lo = make_sw128_kmajor_desc(q_addr + (A::QCH + c) * A::Q_BOX);
#pragma unroll
            for (int k = 0; k < CH / 16; ++k) {
              umma_f16(d, q_lo + 2 * k, k_hi + 2 * k, idesc_s, (c | k) != 0);
              umma_f16(d, q_hi + 2 * k, k_hi + 2 * k, idesc_s, 1);
            }
            release();
            const uint64_t k_lo = take();
#pragma unroll
            for (int k = 0; k < CH / 16; ++k) umma_f16(d, q_hi + 2 * k, k_lo + 2 * k, idesc_s, 1);
            release();
          } else {
#pragma unroll
            for (int k = 0; k < CH / 16; ++k) umma_f16(d, q_hi + 2 * k, k_hi + 2 * k, idesc_s, (c | k) != 0);
            release();
          }
        }
        tcgen05_commit(&s_full[g & 1]);
      };
      for (int w = blockIdx.x; w < total_work; w += gridDim.x) {
        int b, h, q0, len, J;
        decode(w, b, h, q0, len, J);
        if (J == 0) continue;
        mbar_wait(q_bar, wc & 1);
        tcgen05_fence_after();
        issue_s(g0, keys16(0, len));
        if (J == 1) tcgen05_commit(q_free);
        for (int j = 0; j < J; ++j) {
          const int g = g0 + j;
          if (wc == 0) trace_at(p, j, 2, 0);
          if (j + 1 < J) {
            issue_s(g + 1, keys16(j + 1, len));          // the tensor core computes S_{g+1} while the softmax warps work on tile g
            if (j + 2 == J) tcgen05_commit(q_free);       // that was the item's last S product: Q may be overwritten once it completes
          }
          if (wc == 0) trace_at(p, j, 2, 1);
          mbar_wait(&p_full[g & 1], (g >> 1) & 1);
          if (j == 0 && wc > 0) mbar_wait(o_free, (wc - 1) & 1);   // the previous item's epilogue has read O
          tcgen05_fence_after();
          if (wc == 0) trace_at(p, j, 2, 2);
          const uint32_t p_hi = tmem_base + (uint32_t)((g & 1) * BKV), p_lo = p_hi + BKV / 2;   // packed fp16: 64 columns each
          const uint32_t o = tmem_base + A::O_COL;
          const int n16 = keys16(j, len);
          for (int c = 0; c * CH < n16; ++c) {
            const uint64_t v_hi = take();
            const int ks = n16 - c * CH >= CH ? CH / 16 : (n16 - c * CH) / 16;    // K-steps of this box that hold valid keys
            if (X3) {
#pragma unroll
              for (int k = 0; k < CH / 16; ++k) {
                if (k < ks) {
                  umma_f16_ts(o, p_lo + (c * 4 + k) * 8, v_hi + 2 * k, A::IDESC_O, (j | c | k) != 0);
                  umma_f16_ts(o, p_hi + (c * 4 + k) * 8, v_hi + 2 * k, A::IDESC_O, 1);
                }
              }
              release();
              const uint64_t v_lo = take();
#pragma unroll
              for (int k = 0; k < CH / 16; ++k)
                if (k < ks) umma_f16_ts(o, p_hi + (c * 4 + k) * 8, v_lo + 2 * k, A::IDESC_O, 1);
              release();
            } else {
#pragma unroll
              for (int k = 0; k < CH / 16; ++k)
                if (k < ks) umma_f16_ts(o, p_hi + (c * 4 + k) * 8, v_hi + 2 * k, A::IDESC_O, (j | c | k) != 0);
              release();
            }
          }
          tcgen05_commit(&pv_done[0]);                   // lets the softmax warps rescale O if tile g+1 raises the reference max
          if (wc == 0) trace_at(p, j, 2, 3);
        }
        tcgen05_commit(o_full);
        g0 += J; ++wc;
      }
    }
  } else {
    // ---- softmax / epilogue: 8 warps; warps w and w+4 share TMEM lane quarter w%4 and split the columns ----
    const int wq = warp & 3, half = (warp - 2) >> 2;          // half 0: columns [0,64), half 1: [64,128)
    const int row = wq * 32 + lane;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(wq * 32) << 16);
    const float c_exp = p.scale_log2e;
    const long plane = (long)p.B * p.L * p.C;
    float v[64];
    int wc = 0, g0 = 0;
    for (int w = blockIdx.x; w < total_work; w += gridDim.x) {
      int b, h, q0, len, J;
      decode(w, b, h, q0, len, J);
      const int t = q0 + row;
      const bool store = t < p.L;
      const long o_off = ((long)b * p.L + t) * p.C + h * DK + half * (DK / 2);
      if (J == 0) {
        // no valid key at all (len == 0): the reference's masked_fill turns the NaN rows into 0
        if (store) {
          if (p.ctx) for (int c = 0; c < DK / 2; c += 4) *reinterpret_cast<float4*>(p.ctx + o_off + c) = make_float4(0.f, 0.f, 0.f, 0.f);
          if (p.ctxp) for (int c = 0; c < DK / 2; c += 8) {
            *reinterpret_cast<uint4*>(p.ctxp + o_off + c) = make_uint4(0, 0, 0, 0);
            if (X3) *reinterpret_cast<uint4*>(p.ctxp + plane + o_off + c) = make_uint4(0, 0, 0, 0);
          }
        }
        continue;
      }
      float m_ref = -INFINITY, l_row = 0.f;     // reference maximum (raw-score domain) and row sum relative to it
      for (int j = 0; j < J; ++j) {
        const int g = g0 + j;
        if (wq == 0 && wc == 0) trace_at(p, j, half, 0);
        mbar_wait(&s_full[g & 1], (g >> 1) & 1);
        tcgen05_fence_after();
        if (wq == 0 && wc == 0) trace_at(p, j, half, 1);
        const int kv0 = j * BKV + half * 64;
        const bool masked = kv0 + 64 > len;                      // only the last tile of an utterance
        // columns of this half the S product wrote (the MMA warp trims the last tile to the valid keys, rounded to 16):
        // 64 normally; <= 32: half the work; 0: nothing to load, exponentiate or store for this half
        const int ncols = min(max(keys16(j, len) - half * 64, 0), 64);
        __syncwarp();
        const uint32_t tb = lane_addr + (uint32_t)((g & 1) * BKV);
        if (ncols > 32) { tmem_ld32_nowait(tb + half * 64, v); tmem_ld32_nowait(tb + half * 64 + 32, v + 32); tmem_ld_wait_pin<64>(v); }
        else if (ncols > 0) { tmem_ld32_nowait(tb + half * 64, v); tmem_ld_wait_pin<32>(v); }
        if (wq == 0 && wc == 0) trace_at(p, j, half, 2);
        float tmax = -INFINITY;
        if (masked) {
#pragma unroll
          for (int i = 0; i < 64; ++i) if (kv0 + i < len) tmax = fmaxf(tmax, v[i]);     // stale registers past ncols are >= len
        } else {
#pragma unroll
          for (int i = 0; i < 64; ++i) tmax = fmaxf(tmax, v[i]);
        }
        float* xr = xchg + (g & 1) * 2 * BQ;                     // double-buffered exchange: one barrier per tile
        xr[half * BQ + row] = tmax;
        named_bar_sync(1, 256);                                   // also: both halves of every row have read their S columns
        tmax = fmaxf(tmax, xr[(half ^ 1) * BQ + row]);           // both threads of the row now hold the tile's row maximum
        if (wq == 0 && wc == 0) trace_at(p, j, half, 3);
        // lazy reference update: move m_ref only if the tile exceeds it by more than 2^8 in the exp2 domain
        const bool bump = (tmax - m_ref) * c_exp > 8.0f;          // first tile: m_ref = -inf -> always
        if (__any_sync(0xffffffffu, bump)) {
          const float alpha = bump ? fast_exp2((m_ref - tmax) * c_exp) : 1.0f;   // exp2(-inf) = 0 on the first tile
          if (j > 0) {   // O holds tiles 0..j-1: wait until P.V_{g-1} has landed, then scale this thread's half of the row
            mbar_wait(&pv_done[0], (g - 1) & 1);
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
        uint32_t ph[32], pl[32];
        if (ncols > 32) {
          l_row += masked ? softmax_tile<X3, true>(v, c_exp, mb, kv0, len, ph, pl) : softmax_tile<X3, false>(v, c_exp, mb, kv0, len, ph, pl);
          __syncwarp();
          if (wq == 0 && wc == 0) trace_at(p, j, half, 4);
          tmem_st32u(tb + half * 32, ph);                           // P hi: packed columns [0,64) of the tile's buffer
          if (X3) tmem_st32u(tb + BKV / 2 + half * 32, pl);         // P lo: [64,128)
        } else if (ncols > 0) {                                     // short last tile: 32 keys of this half at most
          l_row += softmax_tile<X3, true, 16>(v, c_exp, mb, kv0, len, ph, pl);
          __syncwarp();
          if (wq == 0 && wc == 0) trace_at(p, j, half, 4);
          tmem_st16(tb + half * 32, ph);
          if (X3) tmem_st16(tb + BKV / 2 + half * 32, pl);
        }
        tmem_st_wait();
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_full[g & 1]);
        if (wq == 0 && wc == 0) trace_at(p, j, half, 5);
      }
      named_bar_sync(1, 256);                                     // last exchange buffer is free again
      xchg[half * BQ + row] = l_row;
      named_bar_sync(1, 256);
      l_row += xchg[(half ^ 1) * BQ + row];
      named_bar_sync(1, 256);                                     // both halves have read the sums before the next item's first tile reuses xchg
      // epilogue: O / l -> context rows (0 for masked query rows); each half stores DK/2 columns.  O carries V's
      // kPlaneScale: the planes take it as is, the fp32 rows divide it out.
      mbar_wait(o_full, wc & 1);
      tcgen05_fence_after();
      const float inv = (p.lens && t >= len) ? 0.f : 1.0f / l_row;
#pragma unroll 1
      for (int c0 = 0; c0 < DK / 2; c0 += 32) {
        __syncwarp();
        tmem_ld32(lane_addr + (uint32_t)(A::O_COL + half * (DK / 2) + c0), v);
        if (c0 + 32 >= DK / 2) {                                  // that was this warp's last read of O: the next item may overwrite it
          tcgen05_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(o_free);
        }
        if (!store) continue;
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] *= inv;
        if (p.ctxp != nullptr) {
          uint32_t hh[16], ll[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const float a0 = fminf(fmaxf(v[2 * i], -65504.f), 65504.f), a1 = fminf(fmaxf(v[2 * i + 1], -65504.f), 65504.f);
            const __half2 hv = __floats2half2_rn(a0, a1);
            hh[i] = *reinterpret_cast<const uint32_t*>(&hv);
            if (X3) {
              const float2 gg = __half22float2(hv);
              const __half2 lv = __floats2half2_rn(a0 - gg.x, a1 - gg.y);
              ll[i] = *reinterpret_cast<const uint32_t*>(&lv);
            }
          }
          __half* dh = p.ctxp + o_off + c0;
          st_global_v8_b32(dh, hh); st_global_v8_b32(dh + 16, hh + 8);
          if (X3) { st_global_v8_b32(dh + plane, ll); st_global_v8_b32(dh + plane + 16, ll + 8); }
        }
        if (p.ctx != nullptr) {
          float* dst = p.ctx + o_off + c0;
#pragma unroll
          for (int q = 0; q < 8; ++q)
            *reinterpret_cast<float4*>(dst + q * 4) = make_float4(v[q * 4] * kPlaneInv, v[q * 4 + 1] * kPlaneInv, v[q * 4 + 2] * kPlaneInv, v[q * 4 + 3] * kPlaneInv);
        }
      }
      g0 += J; ++wc;
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  pdl_trigger();                               // FS2_PDL: the next kernel may start launching
  if (warp == 1) {
    tcgen05_fence_after();
    tmem_dealloc(tmem_base, A::TMEM_COLS);
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// Q resident in TMEM (experiment, FS2_ATT_QT=1).  In the kernel above an S = Q K^T instruction streams BOTH operands from
// shared memory (4 KB + 4 KB per 64 cycles at M = N = 128): the clock64 trace shows S issue at ~1.6x its tensor time while
// P.V (A operand = P in TMEM) runs at its tensor time.  Here Q is copied shared -> TMEM once per work item (tcgen05.cp,
// in issue order with the MMAs) and every S product reads only K from shared memory.  TMEM budget with fp32 S / O:
//   Q hi (DK/2 columns) [+ Q lo] | S0 | S1 | O (DK columns)  ->  64 keys per step (3xF16: 192 + 2 x 64 + 192 = 512 exactly).
// Shared memory is one ring of DK x 128-byte slots: Q chunks pass through it like K / V^T tiles.  Persistent CTAs, warp
// roles and softmax as above (two threads per query row, 32 score columns each).
constexpr int BKVQ = 64;

template <int DK, bool X3>
struct QCfg {
  static constexpr int P = X3 ? 2 : 1;
  static constexpr int QCH = DK / CH;
  static constexpr int Q_BOX = BQ * 128;                 // 16 KB: 128 query rows x 64 dk
  static constexpr int K_CHUNK = BKVQ * 128;             // 8 KB: 64 keys x 64 dk
  static constexpr int SLOT = DK * 128;                  // K step = QCH chunks; V^T step = DK rows x 64 keys; one Q chunk fits too
  static constexpr int SLOTS_MAX = (222 * 1024) / SLOT;
  static constexpr int SLOTS = SLOTS_MAX > 12 ? 12 : SLOTS_MAX;
  static constexpr size_t SMEM = (size_t)SLOTS * SLOT + 1024 + 512 + 4 * BQ * 4;
  static constexpr uint32_t IDESC_S = idesc_f16(BQ, BKVQ);
  static constexpr uint32_t IDESC_O = idesc_f16(BQ, DK);
  static constexpr int Q_COLS = DK / 2;                  // packed fp16
  static constexpr int QH_COL = 0, QL_COL = Q_COLS;
  static constexpr int S_COL0 = P * Q_COLS;              // two S/P buffers of 64 columns
  static constexpr int O_COL = S_COL0 + 2 * BKVQ;
  static constexpr int TMEM_COLS = 512;
  static_assert(O_COL + DK <= 512, "TMEM budget");
  static_assert(Q_BOX <= SLOT && QCH * K_CHUNK == SLOT && SLOT % 1024 == 0 && SLOTS >= 6, "ring");
};

template <int DK, bool X3>
__global__ void __launch_bounds__(ATT_THREADS, 1)
attention_qt_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                    const __grid_constant__ CUtensorMap tmap_vt, HParams p) {
  using A = QCfg<DK, X3>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* ring = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(ring + (size_t)A::SLOTS * A::SLOT);
  uint64_t* full_bar = bars;                   // [SLOTS]
  uint64_t* empty_bar = bars + A::SLOTS;       // [SLOTS]
  uint64_t* s_full = bars + 2 * A::SLOTS;      // [2] MMA -> softmax: S tile ready
  uint64_t* pv_done = s_full + 2;              // MMA -> softmax: P.V of the previous step has finished
  uint64_t* p_full = pv_done + 2;              // [2] softmax -> MMA: P written
  uint64_t* o_full = p_full + 2;               // MMA -> epilogue
  uint64_t* o_free = o_full + 1;               // epilogue -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_free + 1);
  float* xchg = reinterpret_cast<float*>(tmem_slot + 4);

  const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;
  const int nq = (p.L + BQ - 1) / BQ;
  const int total_work = p.B * p.heads * nq;
  auto decode = [&](int w, int& b, int& h, int& q0, int& len, int& J) {
    const int qt = w % nq; const int bh = w / nq;
    h = bh % p.heads; b = bh / p.heads; q0 = qt * BQ;
    len = p.lens ? (int)min((long)p.lens[b], (long)p.L) : p.L;
    J = (len + BKVQ - 1) / BKVQ;
  };

  if (threadIdx.x == 0) {
    for (int s = 0; s < A::SLOTS; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&s_full[i], 1); mbar_init(&pv_done[i], 1); mbar_init(&p_full[i], 8); }
    mbar_init(o_full, 1); mbar_init(o_free, 8);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc(tmem_slot, A::TMEM_COLS);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {  // ---- TMA producer: per work item the Q chunks, then K_0, [K_{j+1}, V_j] ... in consumption order ----
      int n = 0;
      auto slot_for = [&](uint32_t bytes) -> uint8_t* {
        const int slot = n % A::SLOTS;
        mbar_wait(&empty_bar[slot], ((n / A::SLOTS) & 1) ^ 1);
        mbar_expect_tx(&full_bar[slot], bytes);
        return ring + (size_t)slot * A::SLOT;
      };
      for (int w = blockIdx.x; w < total_work; w += gridDim.x) {
        int b, h, q0, len, J;
        decode(w, b, h, q0, len, J);
        if (J == 0) continue;
        for (int pl = 0; pl < A::P; ++pl)
          for (int c = 0; c < A::QCH; ++c, ++n) {
            uint8_t* dst = slot_for(A::Q_BOX);
            tma_load_3d(dst, &tmap_q, &full_bar[n % A::SLOTS], h * DK + c * CH, q0, b + pl * p.B);
          }
        auto push_k = [&](int j) {
          for (int pl = 0; pl < A::P; ++pl, ++n) {
            uint8_t* dst = slot_for(A::SLOT);
            for (int c = 0; c < A::QCH; ++c)
              tma_load_3d(dst + (size_t)c * A::K_CHUNK, &tmap_k, &full_bar[n % A::SLOTS], p.C + h * DK + c * CH, j * BKVQ, b + pl * p.B);
          }
        };
        auto push_v = [&](int j) {
          for (int pl = 0; pl < A::P; ++pl, ++n) {
            uint8_t* dst = slot_for(A::SLOT);
            tma_load_3d(dst, &tmap_vt, &full_bar[n % A::SLOTS], j * BKVQ, 0, b * p.heads + h + pl * p.B * p.heads);
          }
        };
        push_k(0);
        for (int j = 0; j < J; ++j) { if (j + 1 < J) push_k(j + 1); push_v(j); }
      }
    }
  } else if (warp == 1) {
    {  // ---- MMA issuer ----
      int n = 0, wc = 0, g0 = 0;
      auto take = [&]() -> uint32_t {
        const int slot = n % A::SLOTS;
        mbar_wait(&full_bar[slot], (n / A::SLOTS) & 1);
        tcgen05_fence_after();
        return smem_u32(ring + (size_t)slot * A::SLOT);
      };
      auto release = [&]() { tcgen05_commit(&empty_bar[n % A::SLOTS]); ++n; };
      const uint32_t qh = tmem_base + A::QH_COL, ql = tmem_base + A::QL_COL;
      auto issue_s = [&](int g) {
        const uint32_t d = tmem_base + (uint32_t)(A::S_COL0 + (g & 1) * BKVQ);
        const uint32_t khi = take();
#pragma unroll
        for (int c = 0; c < A::QCH; ++c) {
          const uint64_t kd = make_sw128_kmajor_desc(khi + c * A::K_CHUNK);
#pragma unroll
          for (int k = 0; k < CH / 16; ++k) {
            if (X3) {
              umma_f16_ts(d, ql + (c * 4 + k) * 8, kd + 2 * k, A::IDESC_S, (c | k) != 0);
              umma_f16_ts(d, qh + (c * 4 + k) * 8, kd + 2 * k, A::IDESC_S, 1);
            } else {
              umma_f16_ts(d, qh + (c * 4 + k) * 8, kd + 2 * k, A::IDESC_S, (c | k) != 0);
            }
          }
        }
        release();
        if (X3) {
          const uint32_t klo = take();
#pragma unroll
          for (int c = 0; c < A::QCH; ++c) {
            const uint64_t kd = make_sw128_kmajor_desc(klo + c * A::K_CHUNK);
#pragma unroll
            for (int k = 0; k < CH / 16; ++k) umma_f16_ts(d, qh + (c * 4 + k) * 8, kd + 2 * k, A::IDESC_S, 1);
          }
          release();
        }
        tcgen05_commit(&s_full[g & 1]);
      };
      for (int w = blockIdx.x; w < total_work; w += gridDim.x) {
        int b, h, q0, len, J;
        decode(w, b, h, q0, len, J);
        if (J == 0) continue;
        // Q: shared -> TMEM, 4 x (128 rows x 32 bytes) per 64-dk chunk; ordered after the previous item's S products by issue order
        for (int pl = 0; pl < A::P; ++pl)
          for (int c = 0; c < A::QCH; ++c) {
            const uint64_t qd = make_sw128_kmajor_desc(take());
#pragma unroll
            for (int k = 0; k < CH / 16; ++k) tmem_cp_128x256b((pl ? ql : qh) + (c * 4 + k) * 8, qd + 2 * k);
            release();
          }
        issue_s(g0);
        for (int j = 0; j < J; ++j) {
          const int g = g0 + j;
          if (wc == 0) trace_at(p, j, 2, 0);
          if (j + 1 < J) issue_s(g + 1);
          if (wc == 0) trace_at(p, j, 2, 1);
          mbar_wait(&p_full[g & 1], (g >> 1) & 1);
          if (j == 0 && wc > 0) mbar_wait(o_free, (wc - 1) & 1);
          tcgen05_fence_after();
          if (wc == 0) trace_at(p, j, 2, 2);
          const uint32_t p_hi = tmem_base + (uint32_t)(A::S_COL0 + (g & 1) * BKVQ), p_lo = p_hi + BKVQ / 2;
          const uint32_t o = tmem_base + A::O_COL;
          const uint64_t vhi = make_sw128_kmajor_desc(take());
#pragma unroll
          for (int k = 0; k < BKVQ / 16; ++k) {
            if (X3) {
              umma_f16_ts(o, p_lo + k * 8, vhi + 2 * k, A::IDESC_O, (j | k) != 0);
              umma_f16_ts(o, p_hi + k * 8, vhi + 2 * k, A::IDESC_O, 1);
            } else {
              umma_f16_ts(o, p_hi + k * 8, vhi + 2 * k, A::IDESC_O, (j | k) != 0);
            }
          }
          release();
          if (X3) {
            const uint64_t vlo = make_sw128_kmajor_desc(take());
#pragma unroll
            for (int k = 0; k < BKVQ / 16; ++k) umma_f16_ts(o, p_hi + k * 8, vlo + 2 * k, A::IDESC_O, 1);
            release();
          }
          tcgen05_commit(&pv_done[0]);
          if (wc == 0) trace_at(p, j, 2, 3);
        }
        tcgen05_commit(o_full);
        g0 += J; ++wc;
      }
    }
  } else {
    // ---- softmax / epilogue: 8 warps; warps w and w+4 share TMEM lane quarter w%4; each thread owns 32 score columns ----
    const int wq = warp & 3, half = (warp - 2) >> 2;
    const int row = wq * 32 + lane;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(wq * 32) << 16);
    const float c_exp = p.scale_log2e;
    const long plane = (long)p.B * p.L * p.C;
    float v[32];
    int wc = 0, g0 = 0;
    for (int w = blockIdx.x; w < total_work; w += gridDim.x) {
      int b, h, q0, len, J;
      decode(w, b, h, q0, len, J);
      const int t = q0 + row;
      const bool store = t < p.L;
      const long o_off = ((long)b * p.L + t) * p.C + h * DK + half * (DK / 2);
      if (J == 0) {
        if (store) {
          if (p.ctx) for (int c = 0; c < DK / 2; c += 4) *reinterpret_cast<float4*>(p.ctx + o_off + c) = make_float4(0.f, 0.f, 0.f, 0.f);
          if (p.ctxp) for (int c = 0; c < DK / 2; c += 8) {
            *reinterpret_cast<uint4*>(p.ctxp + o_off + c) = make_uint4(0, 0, 0, 0);
            if (X3) *reinterpret_cast<uint4*>(p.ctxp + plane + o_off + c) = make_uint4(0, 0, 0, 0);
          }
        }
        continue;
      }
      float m_ref = -INFINITY, l_row = 0.f;
      for (int j = 0; j < J; ++j) {
        const int g = g0 + j;
        if (wq == 0 && wc == 0) trace_at(p, j, half, 0);
        mbar_wait(&s_full[g & 1], (g >> 1) & 1);
        tcgen05_fence_after();
        if (wq == 0 && wc == 0) trace_at(p, j, half, 1);
        const int kv0 = j * BKVQ + half * 32;
        const bool masked = kv0 + 32 > len;
        __syncwarp();
        const uint32_t tb = lane_addr + (uint32_t)(A::S_COL0 + (g & 1) * BKVQ);
        tmem_ld32(tb + half * 32, v);
        if (wq == 0 && wc == 0) trace_at(p, j, half, 2);
        float t0 = -INFINITY, t1 = -INFINITY;
        if (masked) {
#pragma unroll
          for (int i = 0; i < 32; ++i) if (kv0 + i < len) t0 = fmaxf(t0, v[i]);
        } else {
#pragma unroll
          for (int i = 0; i < 16; ++i) { t0 = fmaxf(t0, v[i]); t1 = fmaxf(t1, v[16 + i]); }
        }
        float tmax = fmaxf(t0, t1);
        float* xr = xchg + (g & 1) * 2 * BQ;
        xr[half * BQ + row] = tmax;
        named_bar_sync(1, 256);
        tmax = fmaxf(tmax, xr[(half ^ 1) * BQ + row]);
        if (wq == 0 && wc == 0) trace_at(p, j, half, 3);
        const bool bump = (tmax - m_ref) * c_exp > 8.0f;
        if (__any_sync(0xffffffffu, bump)) {
          const float alpha = bump ? fast_exp2((m_ref - tmax) * c_exp) : 1.0f;
          if (j > 0) {
            mbar_wait(&pv_done[0], (g - 1) & 1);
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
        uint32_t ph[16], pl[16];
        float s0 = 0.f, s1 = 0.f;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          float e0 = fast_exp2(fmaf(v[2 * i], c_exp, -mb)), e1 = fast_exp2(fmaf(v[2 * i + 1], c_exp, -mb));
          if (masked) { if (kv0 + 2 * i >= len) e0 = 0.f; if (kv0 + 2 * i + 1 >= len) e1 = 0.f; }
          const __half2 hh = __floats2half2_rn(e0, e1);
          ph[i] = *reinterpret_cast<const uint32_t*>(&hh);
          const float2 gg = __half22float2(hh);
          if (X3) {
            const __half2 lo2 = __floats2half2_rn(e0 - gg.x, e1 - gg.y);
            pl[i] = *reinterpret_cast<const uint32_t*>(&lo2);
            s0 += e0; s1 += e1;
          } else { s0 += gg.x; s1 += gg.y; }
        }
        l_row += s0 + s1;
        __syncwarp();
        if (wq == 0 && wc == 0) trace_at(p, j, half, 4);
        tmem_st16(tb + half * 16, ph);                          // P hi: packed columns [0,32) of the step's S buffer
        if (X3) tmem_st16(tb + BKVQ / 2 + half * 16, pl);        // P lo: [32,64)
        tmem_st_wait();
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_full[g & 1]);
        if (wq == 0 && wc == 0) trace_at(p, j, half, 5);
      }
      named_bar_sync(1, 256);
      xchg[half * BQ + row] = l_row;
      named_bar_sync(1, 256);
      l_row += xchg[(half ^ 1) * BQ + row];
      named_bar_sync(1, 256);
      mbar_wait(o_full, wc & 1);
      tcgen05_fence_after();
      const float inv = (p.lens && t >= len) ? 0.f : 1.0f / l_row;
#pragma unroll 1
      for (int c0 = 0; c0 < DK / 2; c0 += 32) {
        __syncwarp();
        tmem_ld32(lane_addr + (uint32_t)(A::O_COL + half * (DK / 2) + c0), v);
        if (c0 + 32 >= DK / 2) {
          tcgen05_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(o_free);
        }
        if (!store) continue;
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] *= inv;
        if (p.ctxp != nullptr) {
          uint32_t hh[16], ll[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const float a0 = fminf(fmaxf(v[2 * i], -65504.f), 65504.f), a1 = fminf(fmaxf(v[2 * i + 1], -65504.f), 65504.f);
            const __half2 hv = __floats2half2_rn(a0, a1);
            hh[i] = *reinterpret_cast<const uint32_t*>(&hv);
            if (X3) {
              const float2 gg = __half22float2(hv);
              const __half2 lv = __floats2half2_rn(a0 - gg.x, a1 - gg.y);
              ll[i] = *reinterpret_cast<const uint32_t*>(&lv);
            }
          }
          __half* dh = p.ctxp + o_off + c0;
          st_global_v8_b32(dh, hh); st_global_v8_b32(dh + 16, hh + 8);
          if (X3) { st_global_v8_b32(dh + plane, ll); st_global_v8_b32(dh + plane + 16, ll + 8); }
        }
        if (p.ctx != nullptr) {
          float* dst = p.ctx + o_off + c0;
#pragma unroll
          for (int q = 0; q < 8; ++q)
            *reinterpret_cast<float4*>(dst + q * 4) = make_float4(v[q * 4] * kPlaneInv, v[q * 4 + 1] * kPlaneInv, v[q * 4 + 2] * kPlaneInv, v[q * 4 + 3] * kPlaneInv);
        }
      }
      g0 += J; ++wc;
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    tmem_dealloc(tmem_base, A::TMEM_COLS);
  }
}

template <int DK, bool X3>
int launch_qt(const __half* qkp, const __half* vtp, int lpad, const int64_t* lens, int B, int L, int C, int heads, float* ctx,
              __half* ctxp, cudaStream_t st) {
  using A = QCfg<DK, X3>;
  static unsigned long long configured = 0;
  int rc;
  if ((rc = ensure_smem_attr(attention_qt_kernel<DK, X3>, A::SMEM, &configured))) return rc;
  CUtensorMap mq, mk, mvt;
  const uint64_t row = (uint64_t)2 * C * 2;
  if ((rc = make_map(&mq, qkp, (uint64_t)2 * C, L, (uint64_t)B * A::P, row, row * L, BQ, true))) return rc;
  if ((rc = make_map(&mk, qkp, (uint64_t)2 * C, L, (uint64_t)B * A::P, row, row * L, BKVQ, true))) return rc;
  if ((rc = make_map(&mvt, vtp, L, DK, (uint64_t)B * heads * A::P, (uint64_t)lpad * 2, (uint64_t)lpad * 2 * DK, DK, true))) return rc;
  HParams p;
  p.lens = lens; p.B = B; p.L = L; p.C = C; p.heads = heads; p.ctx = ctx; p.ctxp = ctxp;
  p.scale_log2e = (1.0f / sqrtf((float)DK)) * 1.4426950408889634f * kPlaneInv * kPlaneInv;
  p.debug = 0;
  p.trace = att_trace_buffer();
  const long work = (long)B * heads * ((L + BQ - 1) / BQ);
  const int grid = (int)(work < sm_count_current() ? work : sm_count_current());
  attention_qt_kernel<DK, X3><<<grid, ATT_THREADS, A::SMEM, st>>>(mq, mk, mvt, p);
  FS2_LAUNCH_CHECK();
  if (p.trace) att_trace_dump(p.trace, (L + BKVQ - 1) / BKVQ, st);
  return FS2_OK;
}

// ---------------------------------------------------------------------------------------------------------------------
// F16, two query tiles in flight per CTA (FS2_MATH_F16's decoder): the single-tile kernel above leaves the tensor pipe idle
// while a tile's softmax runs (S -> exp2 -> P -> P.V is one dependency chain per tile; profiles/r01_ncu_attn_tf32_v18.md:
// tensor pipe 33 % active).  Here a CTA owns TWO 128-query tiles A and B of the same (batch, head), each with its own S/P
// buffer, O accumulator and softmax warpgroup, sharing one K / V^T stream (half the L2 traffic per query): the MMA warp
// alternates   P.V_A(j), S_A(j+1) | P.V_B(j), S_B(j+1)   so one tile's MMAs run under the other tile's softmax.
//   keys per step: 64 -> S tile 128 x 64 fp32 = 64 TMEM columns; TMEM = S_A | S_B | O_A | O_B = 64 + 64 + 2 DK <= 512
//   one softmax thread owns a whole 64-score row: no cross-thread exchange, no block barrier in the loop
//   shared memory: Q_A, Q_B resident (2 x DK/64 x 16 KB) + a ring of 64-key K boxes / V^T boxes (DK x 128 B each)
constexpr int BKV2 = 64;
constexpr int ATT2_THREADS = 320;   // TMA, MMA, 4 softmax warps per tile

template <int DK>
struct H2Cfg {
  static constexpr int QCH = DK / CH;
  static constexpr int Q_BOX = BQ * 128;                 // 16 KB: 128 query rows x 64 dk
  static constexpr int Q_TILE = QCH * Q_BOX;
  static constexpr int Q_BYTES = 2 * Q_TILE;
  static constexpr int K_CHUNK = BKV2 * 128;             // 8 KB: 64 keys x 64 dk
  static constexpr int SLOT = DK * 128;                  // K step = QCH chunks = DK * 128 B; V^T step = DK rows x 64 keys: same size
  static constexpr int SLOTS_MAX = (222 * 1024 - Q_BYTES) / SLOT;
  static constexpr int SLOTS = SLOTS_MAX > 8 ? 8 : SLOTS_MAX;
  static constexpr size_t SMEM = (size_t)Q_BYTES + (size_t)SLOTS * SLOT + 1024 + 512;
  static constexpr uint32_t IDESC_S = idesc_f16(BQ, BKV2);
  static constexpr uint32_t IDESC_O = idesc_f16(BQ, DK);
  static constexpr int TMEM_COLS = 512;
  static constexpr int S_COL(int x) { return x * BKV2; }
  static constexpr int O_COL(int x) { return 2 * BKV2 + x * DK; }
  static_assert(DK % 64 == 0 && 2 * BKV2 + 2 * DK <= 512, "TMEM budget");
  static_assert(QCH * K_CHUNK == SLOT && SLOT % 1024 == 0 && SLOTS >= 4, "ring");
};

template <int DK>
__global__ void __launch_bounds__(ATT2_THREADS, 1)
attention_f16x2_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                       const __grid_constant__ CUtensorMap tmap_vt, HParams p) {
  using A = H2Cfg<DK>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* q_smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* ring = q_smem + A::Q_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(ring + (size_t)A::SLOTS * A::SLOT);
  uint64_t* full_bar = bars;                   // [SLOTS]
  uint64_t* empty_bar = bars + A::SLOTS;       // [SLOTS]
  uint64_t* q_bar = bars + 2 * A::SLOTS;
  uint64_t* s_full = q_bar + 1;                // [2] MMA -> softmax X: S tile ready
  uint64_t* p_full = s_full + 2;               // [2] softmax X -> MMA: P written
  uint64_t* pv_done = p_full + 2;              // [2] MMA -> softmax X: P.V of the previous step has finished (O may be rescaled)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(pv_done + 2);

  const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * 2 * BQ, h = blockIdx.y, b = blockIdx.z;
  const int len = p.lens ? (int)min((long)p.lens[b], (long)p.L) : p.L;
  const int J = (len + BKV2 - 1) / BKV2;
  const bool has_b = q0 + BQ < p.L;            // the second tile exists (CTA-uniform)

  if (threadIdx.x == 0) {
    for (int s = 0; s < A::SLOTS; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(q_bar, 1);
    for (int i = 0; i < 2; ++i) { mbar_init(&s_full[i], 1); mbar_init(&p_full[i], 4); mbar_init(&pv_done[i], 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc(tmem_slot, A::TMEM_COLS);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (J > 0) {
    if (warp == 0) {
      if (lane == 0) {  // ---- TMA producer: both Q tiles once, then K_0, V_0, K_1, V_1, ... ----
        mbar_expect_tx(q_bar, has_b ? A::Q_BYTES : A::Q_TILE);
        for (int x = 0; x < (has_b ? 2 : 1); ++x)
          for (int c = 0; c < A::QCH; ++c)
            tma_load_3d(q_smem + (size_t)x * A::Q_TILE + (size_t)c * A::Q_BOX, &tmap_q, q_bar, h * DK + c * CH, q0 + x * BQ, b);
        for (int n = 0; n < 2 * J; ++n) {
          const int slot = n % A::SLOTS, j = n >> 1;
          mbar_wait(&empty_bar[slot], ((n / A::SLOTS) & 1) ^ 1);
          mbar_expect_tx(&full_bar[slot], A::SLOT);
          uint8_t* dst = ring + (size_t)slot * A::SLOT;
          if (n & 1) {
            tma_load_3d(dst, &tmap_vt, &full_bar[slot], j * BKV2, 0, b * p.heads + h);
          } else {
            for (int c = 0; c < A::QCH; ++c)
              tma_load_3d(dst + (size_t)c * A::K_CHUNK, &tmap_k, &full_bar[slot], p.C + h * DK + c * CH, j * BKV2, b);
          }
        }
      }
    } else if (warp == 1) {
      {  // ---- MMA issuer (warp-uniform loop, one lane elected inside each tcgen05 asm) ----
        mbar_wait(q_bar, 0);
        tcgen05_fence_after();
        const uint32_t q_addr = smem_u32(q_smem);
        auto slot_ready = [&](int n) -> uint32_t {
          const int slot = n % A::SLOTS;
          mbar_wait(&full_bar[slot], (n / A::SLOTS) & 1);
          tcgen05_fence_after();
          return smem_u32(ring + (size_t)slot * A::SLOT);
        };
        auto issue_s = [&](int x, int j) {      // S_x(j) = Q_x K_j^T -> S/P buffer x
          const uint32_t kbase = slot_ready(2 * j);
          const uint32_t d = tmem_base + (uint32_t)A::S_COL(x);
#pragma unroll
          for (int c = 0; c < A::QCH; ++c) {
            const uint64_t qd = make_sw128_kmajor_desc(q_addr + x * A::Q_TILE + c * A::Q_BOX);
            const uint64_t kd = make_sw128_kmajor_desc(kbase + c * A::K_CHUNK);
#pragma unroll
            for (int k = 0; k < CH / 16; ++k) umma_f16(d, qd + 2 * k, kd + 2 * k, A::IDESC_S, (c | k) != 0);
          }
          if (x == 1 || !has_b) tcgen05_commit(&empty_bar[(2 * j) % A::SLOTS]);   // last reader of K_j
          tcgen05_commit(&s_full[x]);
        };
        auto issue_pv = [&](int x, int j) {     // O_x += P_x(j) V_j
          mbar_wait(&p_full[x], j & 1);
          tcgen05_fence_after();
          const uint64_t vd = make_sw128_kmajor_desc(slot_ready(2 * j + 1));
          const uint32_t pa = tmem_base + (uint32_t)A::S_COL(x), o = tmem_base + (uint32_t)A::O_COL(x);
#pragma unroll
          for (int k = 0; k < BKV2 / 16; ++k) umma_f16_ts(o, pa + k * 8, vd + 2 * k, A::IDESC_O, (j | k) != 0);
          if (x == 1 || !has_b) tcgen05_commit(&empty_bar[(2 * j + 1) % A::SLOTS]);   // last reader of V_j
          tcgen05_commit(&pv_done[x]);
        };
        issue_s(0, 0);
        if (has_b) issue_s(1, 0);
        for (int j = 0; j < J; ++j) {
          issue_pv(0, j);
          if (j + 1 < J) issue_s(0, j + 1);
          if (has_b) {
            issue_pv(1, j);
            if (j + 1 < J) issue_s(1, j + 1);
          }
        }
      }
    } else {
      // ---- softmax / epilogue: warps 2-5 = tile A, 6-9 = tile B; thread == query row == TMEM lane ----
      const int x = (warp - 2) >> 2, wq = warp & 3;
      if (x == 0 || has_b) {
        const int row = wq * 32 + lane;
        const uint32_t lane_addr = tmem_base + ((uint32_t)(wq * 32) << 16);
        const uint32_t s_addr = lane_addr + (uint32_t)A::S_COL(x), o_addr = lane_addr + (uint32_t)A::O_COL(x);
        float v[64];
        float m_ref = -INFINITY, l_row = 0.f;
        const float c_exp = p.scale_log2e;
        for (int j = 0; j < J; ++j) {
          mbar_wait(&s_full[x], j & 1);
          tcgen05_fence_after();
          const int kv0 = j * BKV2;
          const bool masked = kv0 + BKV2 > len;
          __syncwarp();
          if (!(p.debug & 2)) { tmem_ld32_nowait(s_addr, v); tmem_ld32_nowait(s_addr + 32, v + 32); tmem_ld_wait_pin<64>(v); }
          else {
#pragma unroll
            for (int i = 0; i < 64; ++i) v[i] = (float)(i + row) * 1e-3f;
          }
          float t0 = -INFINITY, t1 = -INFINITY, t2 = -INFINITY, t3 = -INFINITY;     // four chains: the serial max is latency
          if (masked) {
#pragma unroll
            for (int i = 0; i < 64; ++i) if (kv0 + i < len) t0 = fmaxf(t0, v[i]);
          } else {
#pragma unroll
            for (int i = 0; i < 16; ++i) { t0 = fmaxf(t0, v[i]); t1 = fmaxf(t1, v[16 + i]); t2 = fmaxf(t2, v[32 + i]); t3 = fmaxf(t3, v[48 + i]); }
          }
          const float tmax = fmaxf(fmaxf(t0, t1), fmaxf(t2, t3));
          const bool bump = (tmax - m_ref) * c_exp > 8.0f;
          if (__any_sync(0xffffffffu, bump)) {
            const float alpha = bump ? fast_exp2((m_ref - tmax) * c_exp) : 1.0f;
            if (j > 0) {
              mbar_wait(&pv_done[x], (j - 1) & 1);
              tcgen05_fence_after();
              float o[32];
#pragma unroll 1
              for (int c0 = 0; c0 < DK; c0 += 32) {
                __syncwarp();
                tmem_ld32(o_addr + c0, o);
#pragma unroll
                for (int i = 0; i < 32; ++i) o[i] *= alpha;
                tmem_st32(o_addr + c0, o);
              }
              tmem_st_wait();
            }
            l_row *= alpha;
            if (bump) m_ref = tmax;
          }
          const float mb = m_ref * c_exp;
          uint32_t ph[32], pl_unused[1];
          if (!(p.debug & 1)) {
            l_row += masked ? softmax_tile<false, true>(v, c_exp, mb, kv0, len, ph, pl_unused) : softmax_tile<false, false>(v, c_exp, mb, kv0, len, ph, pl_unused);
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i) ph[i] = __float_as_uint(v[2 * i]);
            l_row += 1.f;
          }
          __syncwarp();
          if (!(p.debug & 4)) {
            tmem_st32u(s_addr, ph);                                 // P (packed fp16) over the first 32 columns of the S tile
            tmem_st_wait();
          }
          tcgen05_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&p_full[x]);
        }
        // epilogue: O / l -> hi plane of the context (and / or fp32 rows)
        mbar_wait(&pv_done[x], (J - 1) & 1);
        tcgen05_fence_after();
        const int t = q0 + x * BQ + row;
        const bool store = t < p.L;
        const float inv = (p.lens && t >= len) ? 0.f : 1.0f / l_row;
        const long o_off = ((long)b * p.L + t) * p.C + h * DK;
#pragma unroll 1
        for (int c0 = 0; c0 < DK; c0 += 32) {
          __syncwarp();
          tmem_ld32(o_addr + c0, v);
          if (!store) continue;
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] *= inv;
          if (p.ctxp != nullptr) {
            uint32_t hh[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const __half2 hv = __floats2half2_rn(fminf(fmaxf(v[2 * i], -65504.f), 65504.f), fminf(fmaxf(v[2 * i + 1], -65504.f), 65504.f));
              hh[i] = *reinterpret_cast<const uint32_t*>(&hv);
            }
            st_global_v8_b32(p.ctxp + o_off + c0, hh); st_global_v8_b32(p.ctxp + o_off + c0 + 16, hh + 8);
          }
          if (p.ctx != nullptr) {
            float* dst = p.ctx + o_off + c0;
#pragma unroll
            for (int q = 0; q < 8; ++q)
              *reinterpret_cast<float4*>(dst + q * 4) = make_float4(v[q * 4] * kPlaneInv, v[q * 4 + 1] * kPlaneInv, v[q * 4 + 2] * kPlaneInv, v[q * 4 + 3] * kPlaneInv);
          }
        }
      }
    }
  } else if (warp >= 2) {
    // no valid key at all (len == 0): the reference's masked_fill turns the NaN rows into 0
    const int x = (warp - 2) >> 2;
    const int t = q0 + x * BQ + (warp & 3) * 32 + lane;
    if (t < p.L) {
      const long o_off = ((long)b * p.L + t) * p.C + h * DK;
      if (p.ctx) for (int c = 0; c < DK; c += 4) *reinterpret_cast<float4*>(p.ctx + o_off + c) = make_float4(0.f, 0.f, 0.f, 0.f);
      if (p.ctxp) for (int c = 0; c < DK; c += 8) *reinterpret_cast<uint4*>(p.ctxp + o_off + c) = make_uint4(0, 0, 0, 0);
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
int launch_x2(const __half* qkp, const __half* vtp, int lpad, const int64_t* lens, int B, int L, int C, int heads, float* ctx,
              __half* ctxp, cudaStream_t st) {
  using A = H2Cfg<DK>;
  static unsigned long long configured = 0;   // per-device bit mask
  int rc;
  if ((rc = ensure_smem_attr(attention_f16x2_kernel<DK>, A::SMEM, &configured))) return rc;
  CUtensorMap mq, mk, mvt;
  const uint64_t row = (uint64_t)2 * C * 2;
  if ((rc = make_map(&mq, qkp, (uint64_t)2 * C, L, B, row, row * L, BQ, true))) return rc;
  if ((rc = make_map(&mk, qkp, (uint64_t)2 * C, L, B, row, row * L, BKV2, true))) return rc;
  if ((rc = make_map(&mvt, vtp, L, DK, (uint64_t)B * heads, (uint64_t)lpad * 2, (uint64_t)lpad * 2 * DK, DK, true))) return rc;
  HParams p;
  p.lens = lens; p.B = B; p.L = L; p.C = C; p.heads = heads; p.ctx = ctx; p.ctxp = ctxp;
  p.scale_log2e = (1.0f / sqrtf((float)DK)) * 1.4426950408889634f * kPlaneInv * kPlaneInv;
  p.debug = att_debug();
  p.trace = nullptr;
  dim3 grid((L + 2 * BQ - 1) / (2 * BQ), heads, B);
  attention_f16x2_kernel<DK><<<grid, ATT2_THREADS, A::SMEM, st>>>(mq, mk, mvt, p);
  FS2_LAUNCH_CHECK();
  return FS2_OK;
}

template <int DK, bool X3>
int launch(const __half* qkp, const __half* vtp, int lpad, const int64_t* lens, int B, int L, int C, int heads, float* ctx,
           __half* ctxp, cudaStream_t st) {
  using A = HCfg<DK, X3>;
  static unsigned long long configured = 0;   // per-device bit mask
  int rc;
  if ((rc = ensure_smem_attr(attention_f16_kernel<DK, X3>, A::SMEM, &configured))) return rc;
  CUtensorMap mqk, mvt;
  const uint64_t row = (uint64_t)2 * C * 2;
  if ((rc = make_map(&mqk, qkp, (uint64_t)2 * C, L, (uint64_t)B * A::P, row, row * L, BQ, true))) return rc;
  // extent L (not lpad) along kv: alignment padding columns are never read, TMA zero-fills past L
  if ((rc = make_map(&mvt, vtp, L, DK, (uint64_t)B * heads * A::P, (uint64_t)lpad * 2, (uint64_t)lpad * 2 * DK, DK, true))) return rc;
  HParams p;
  p.lens = lens; p.B = B; p.L = L; p.C = C; p.heads = heads; p.ctx = ctx; p.ctxp = ctxp;
  p.scale_log2e = (1.0f / sqrtf((float)DK)) * 1.4426950408889634f * kPlaneInv * kPlaneInv;
  p.debug = 0;
  p.trace = att_trace_buffer();
  const long work = (long)B * heads * ((L + BQ - 1) / BQ);
  const int grid = (int)(work < sm_count_current() ? work : sm_count_current());     // persistent: one CTA per SM
  FS2_CUDA_CHECK(launch_pdl(attention_f16_kernel<DK, X3>, dim3(grid), dim3(ATT_THREADS), A::SMEM, st, mqk, mvt, p));
  FS2_LAUNCH_CHECK();
  if (p.trace) att_trace_dump(p.trace, (L + BKV - 1) / BKV, st);
  return FS2_OK;
}

// test helper (single-operator entry): qkv fp32 [B,L,3C] -> q|k planes [2][B*L][2C] and V^T planes [2][B*heads][dk][lpad]
__global__ void qkv_to_planes_kernel(const float* __restrict__ qkv, int B, int L, int C, __half* __restrict__ qkp,
                                     __half* __restrict__ vtp, int lpad) {
  const long rows = (long)B * L;
  const long total = rows * 3 * C;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const long r = i / (3 * C); const int c = (int)(i - r * 3 * C);
    const float x = fminf(fmaxf(qkv[i] * kPlaneScale, -65504.f), 65504.f);
    const __half hi = __float2half_rn(x), lo = __float2half_rn(x - __half2float(hi));
    if (c < 2 * C) {
      qkp[r * 2 * C + c] = hi; qkp[(rows + r) * 2 * C + c] = lo;
    } else {
      const long bb = r / L; const int t = (int)(r - bb * L);
      const long o = (bb * C + (c - 2 * C)) * (long)lpad + t;           // (b*heads + h)*dk + d == b*C + n
      vtp[o] = hi; vtp[(long)B * C * lpad + o] = lo;
    }
  }
}

}  // namespace

int qkv_to_planes(const float* qkv, int B, int L, int C, int heads, __half* qkp, __half* vtp, int lpad, cudaStream_t st) {
  (void)heads;
  const long total = (long)B * L * 3 * C;
  if (total == 0) return FS2_OK;
  long blocks = (total + 255) / 256;
  qkv_to_planes_kernel<<<(int)(blocks > 148 * 16 ? 148 * 16 : blocks), 256, 0, st>>>(qkv, B, L, C, qkp, vtp, lpad);
  FS2_LAUNCH_CHECK();
  return FS2_OK;
}

int attention_planes(const __half* qkp, const __half* vtp, int lpad, const int64_t* lens, int B, int L, int C, int heads, bool x3,
                     float* ctx, __half* ctxp, cudaStream_t st) {
  FS2_REQUIRE(heads > 0 && C % heads == 0, "attention: C=%d not divisible by heads=%d", C, heads);
  FS2_REQUIRE(qkp && vtp && lpad >= L && lpad % 8 == 0, "attention_planes: needs q|k planes and transposed V planes with a 16-byte aligned row pitch");
  FS2_REQUIRE(ctx || ctxp, "attention_planes: no output");
  FS2_REQUIRE(!ctxp || ((reinterpret_cast<uintptr_t>(ctxp) & 31) == 0 && (((long)B * L * C) % 16) == 0), "attention_planes: context planes must be 32-byte aligned");
  if (B == 0 || L == 0) return FS2_OK;
  const int dk = C / heads;
  {
    static int use_qt = -1;   // FS2_ATT_QT=1 (experiment): Q resident in TMEM, 64 keys per step
    if (use_qt < 0) { const char* e = getenv("FS2_ATT_QT"); use_qt = e ? atoi(e) : 0; }
    if (use_qt && dk == 192) return x3 ? launch_qt<192, true>(qkp, vtp, lpad, lens, B, L, C, heads, ctx, ctxp, st) : launch_qt<192, false>(qkp, vtp, lpad, lens, B, L, C, heads, ctx, ctxp, st);
    if (use_qt && dk == 128) return x3 ? launch_qt<128, true>(qkp, vtp, lpad, lens, B, L, C, heads, ctx, ctxp, st) : launch_qt<128, false>(qkp, vtp, lpad, lens, B, L, C, heads, ctx, ctxp, st);
  }
  if (!x3) {
    // FS2_ATT_X2=1 (experiment): the two-tile kernel.  Measured slower than the single-tile one (c2: 0.50 vs 0.42 ms per
    // step, c4: 1.09 vs 0.77, profiles/r02_attention_ab.md): with 64 keys per step it pays twice the softmax <-> MMA
    // hand-offs per key, and the softmax itself is bound by the TMEM read of S and MUFU.EX2, not by MMA overlap
    static int use_x2 = -1;
    if (use_x2 < 0) { const char* e = getenv("FS2_ATT_X2"); use_x2 = e ? atoi(e) : 0; }
    if (use_x2 && dk == 192) return launch_x2<192>(qkp, vtp, lpad, lens, B, L, C, heads, ctx, ctxp, st);
    if (use_x2 && dk == 128) return launch_x2<128>(qkp, vtp, lpad, lens, B, L, C, heads, ctx, ctxp, st);
  }
  if (dk == 192) return x3 ? launch<192, true>(qkp, vtp, lpad, lens, B, L, C, heads, ctx, ctxp, st) : launch<192, false>(qkp, vtp, lpad, lens, B, L, C, heads, ctx, ctxp, st);
  if (dk == 128) return x3 ? launch<128, true>(qkp, vtp, lpad, lens, B, L, C, heads, ctx, ctxp, st) : launch<128, false>(qkp, vtp, lpad, lens, B, L, C, heads, ctx, ctxp, st);
  set_error("attention_planes: d_k=%d unsupported (128 or 192)", dk);
  return FS2_ERR_INVALID;
}

}  // namespace fs2
