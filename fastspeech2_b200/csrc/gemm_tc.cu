// Tensor-core "tap GEMM" for sm_100a: tcgen05.mma (kind::tf32 or kind::f16) with TMEM accumulators,
// operands staged in shared memory by TMA (cp.async.bulk.tensor, 128-byte swizzle), mbarrier
// pipelines, persistent CTAs.
//
//   out[b,t,n] = act( sum_{j<taps} sum_{k<K} x[b, t+j-pad, k] * w[j][n][k] + bias[n] ) (+ resid[b,t,n])
//
// Same contract as gemm_fp32.cu (the CUDA-core family).  Three instantiation families:
//   <BN, false, false>  plain tf32 (one MMA per product) on the fp32 activations themselves: the decoder side in
//                       FS2_MATH_TF32, and the GEMMs whose A operand has no fp16 copy in FS2_MATH_F16;
//   <BN, false, true>   kind::f16 on fp16 copies of activations and weights (FS2_MATH_F16: q|k|v, conv-FFN, mel
//                       projection, Postnet); see "HALF" below;
//   <BN, true,  true>   error-compensated "3xF16" (encoder and predictors in every tensor-core mode, everything in
//                       FS2_MATH_3XTF32): their outputs feed round() / bucketize(), where 10-bit-mantissa noise
//                       (~1e-3) would flip integers; see "PRECISE && HALF" below;
//
// Why no im2col: activations are [B, time, channel] fp32 with channels innermost, which *is* the
// K-major A operand of a GEMM.  Tap j of a 1-D convolution is the same matrix shifted by
// (j - pad) rows, so the producer just issues the TMA box at row coordinate t0 + j - pad of a
// 3-D tensor map {channel, time, utterance}; rows outside [0, L) of the utterance are
// zero-filled by the TMA unit (that is exactly Conv1d's "same" padding), and fp32 data in shared
// memory is consumed directly by kind::tf32, so there is no conversion pass either.
// K loop = taps x ceil(K / 32) pipeline steps of 4 (12 when PRECISE) MMAs with K = 8 each (tf32; f16: ceil(K / 64)
// steps of the same MMA count with K = 16 each).
//
// One persistent CTA per SM walks the 128 x BN output tiles (n fastest, so concurrently running
// CTAs share weight tiles in L2).  Three pipelines:
//   smem ring   : warp 0 (TMA producer, one lane)  <-> warp 1 (MMA issuer, one lane), full/empty mbarriers
//   TMEM        : 2 (BN > 128) or 4 accumulator buffers; warp 1 fills buffer i % NACC while the epilogue drains older ones
//   output      : two epilogue groups (warps 2-5 and 6-9; thread == output row == TMEM lane) take alternate
//                 32-column chunks: tcgen05.ld -> bias / ReLU / tanh / residual in registers -> four
//                 256-bit global stores per thread (sector-complete, no shared-memory transpose, no
//                 barriers; rows past the utterance end are predicated off)
// Convolutions tile each utterance separately so the shifted boxes never cross an utterance boundary:
// floor(L/128) full row tiles per utterance, and the tails (L % 128 rows, in 16-row granules loaded by
// separate small TMA boxes) of several utterances packed into shared tiles, so no tensor-core rows are
// spent on padding (at L = 800 that was 12 %).  Plain GEMMs (taps == 1) tile the flat [B*L, K] matrix.
// Every mbarrier wait is bounded: a pipeline bug traps instead of hanging the GPU.
//
// PRECISE && HALF ("3xF16", the default error-compensated family): a small pre-pass (split_rows_f16_kernel) writes the
// fp32 activations as two fp16 planes, hi = rn(x) and lo = rn(x - hi), the weights are split the same way at load
// time, and the GEMM loads all four operand tiles by TMA and runs the three products on kind::f16 -- no split warps, no
// shared-memory rewrite, and a pipeline step covers twice the K for the same 12 MMAs.  hi + lo carries 22 mantissa
// bits where lo is a normal fp16 (|x| >= 2^-3); smaller elements -- all weights of a trained layer -- keep an absolute
// error of up to 2^-25 = 3e-8 (subnormal lo), an output error floor of ~sqrt(K) * rms(x) * 2e-8 that sits well below
// the tensor core's own accumulation error for O(1) results (tests/test_numerics_model.py).  The lo plane is addressed through the same tensor
// map: plane stride = B*L rows, i.e. utterance index b + B.
//
// HALF = true, PRECISE = false (FS2_MATH_F16, the decoder's conv-FFN): the same pipeline on fp16 copies of the activations and weights
// with kind::f16 -- a 128-byte swizzle row then holds 64 K-elements and one MMA covers K = 16, so a pipeline step moves
// the same bytes and issues the same four MMAs but does twice the work; the epilogue can emit the result as fp16 for
// the next f16 GEMM (conv k=9 -> ReLU -> conv k=1).  fp16 has tf32's 10-bit mantissa; accumulation stays fp32.
#include <stdlib.h>

#include "tc_common.cuh"

namespace fs2 {
namespace {
using namespace tc;

constexpr int BM = 128;
constexpr int BK = 32;                 // fp32 elements per pipeline step = one 128-byte swizzle row (64 when the operands are fp16)
constexpr int UMMA_K = 8;              // tf32 (16 for f16: 32 bytes of K per instruction either way)
constexpr int A_BYTES = BM * BK * 4;   // 16 KB
constexpr int STAGING_BYTES = 8 * 1024;          // bias[N] (N <= 2048) staged once per CTA; outputs go straight from registers
constexpr int RING_BUDGET = 227 * 1024 - STAGING_BYTES - 1024 /*align slack*/ - 512 /*barriers*/;

struct TcParams {
  int L, tiles_per_utt;          // tiles_per_utt == 0: flat tiling, L = B*L rows
  int m_tiles, n_tiles;
  // convolutions (tiles_per_utt > 0): every utterance has `full` 128-row tiles; its tail (L % 128 rows, `gn` granules
  // of 16 rows) shares a packed tile with the tails of upt - 1 other utterances, so no MMA rows are wasted on padding
  int B, full, gn, upt, full_tiles;
  int K, taps, pad;
  const float* bias; const float* resid; int ldr; int act;
  float* out; int ldo;
  __half* out_h; int ldo_h;      // HALF only: fp16 copy of the result (out may then be null)
  // optional: columns >= vt_col0 are the V third of a q|k|v projection and are stored transposed,
  // vt[(b*heads + h)*dk + d][t] with row pitch vt_lpad, for the attention kernel's K-major P.V operand
  float* vt_out; int vt_col0, vt_dk, vt_heads, vt_lpad, vt_L;
  int debug;   // FS2_GEMM_DEBUG: bit0 skip tcgen05.ld, bit1 skip stores (profiling experiments only)
};

constexpr int pow2_at_least(int x) { return x <= 32 ? 32 : x <= 64 ? 64 : x <= 128 ? 128 : 256; }

template <int BN, bool PRECISE, bool HALF = false>
struct Cfg {
  static constexpr int B_BYTES = BN * BK * 4;
  static constexpr int STAGE_BYTES = (PRECISE ? 2 : 1) * (A_BYTES + B_BYTES);   // [A(hi)][A lo][B hi][B lo]
  static constexpr bool SPLIT16 = PRECISE;                // A hi / lo are fp16 planes in global memory, loaded like B hi / lo
  static_assert(!PRECISE || HALF, "the error-compensated family is 3xF16");
  static constexpr int STAGES = (RING_BUDGET / STAGE_BYTES) > 8 ? 8 : (RING_BUDGET / STAGE_BYTES);
  static constexpr int ACC_STRIDE = pow2_at_least(BN);     // TMEM columns per accumulator buffer
  static constexpr int NACC = 512 / ACC_STRIDE > 4 ? 4 : 512 / ACC_STRIDE;   // accumulator buffers in flight (2 for BN > 128, else 4)
  static constexpr int TMEM_COLS = NACC * ACC_STRIDE;
  static constexpr int GROUPS = 2;       // epilogue warp groups (4 warps each), alternate 32-column chunks
  static constexpr int THREADS = 64 + GROUPS * 128;
  static constexpr size_t SMEM = (size_t)STAGES * STAGE_BYTES + STAGING_BYTES + 1024 + 512;
  static constexpr uint32_t IDESC = HALF ? idesc_f16(BM, BN) : idesc_tf32(BM, BN);
  static constexpr int BKE = HALF ? 2 * BK : BK;           // K elements per pipeline step

  static constexpr int A_LO = A_BYTES;                                  // offsets inside a stage
  static constexpr int B_HI = PRECISE ? 2 * A_BYTES : A_BYTES;
  static constexpr int B_LO = B_HI + B_BYTES;
  static constexpr uint32_t TX_BYTES = (SPLIT16 ? 2 : 1) * A_BYTES + (PRECISE ? 2 : 1) * B_BYTES;
  static_assert(BN % 16 == 0 && BN >= 16 && BN <= 256, "UMMA N for M=128");
  static_assert(B_BYTES % 1024 == 0, "B stage must keep 1024-byte alignment");
  static_assert(STAGES >= 2 && TMEM_COLS <= 512, "resources");
};

template <int BN, bool PRECISE, bool HALF>
__global__ void __launch_bounds__(Cfg<BN, PRECISE, HALF>::THREADS, 1)
tap_gemm_tf32_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                     const __grid_constant__ CUtensorMap tmap_b_lo, const __grid_constant__ CUtensorMap tmap_a16, TcParams p) {
  using C = Cfg<BN, PRECISE, HALF>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* tiles = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // keeps the shared address space (no generic LD/ST)
  uint8_t* staging = tiles + (size_t)C::STAGES * C::STAGE_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(staging + STAGING_BYTES);
  uint64_t* empty_bar = full_bar + C::STAGES;
  uint64_t* acc_full = empty_bar + C::STAGES;     // [NACC] MMA -> epilogue
  uint64_t* acc_empty = acc_full + 4;             // [NACC] epilogue -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int kchunks = (p.K + C::BKE - 1) / C::BKE;
  const int steps = p.taps * kchunks;
  const int total_tiles = p.m_tiles * p.n_tiles;

  if (threadIdx.x == 0) {
    for (int s = 0; s < C::STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int i = 0; i < C::NACC; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], 4 * C::GROUPS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc(tmem_slot, C::TMEM_COLS);   // whole warp: both accumulator buffers
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  float* bias_s = reinterpret_cast<float*>(staging);      // whole bias vector, read back as broadcast LDS in the epilogue
  {
    const int N = p.n_tiles * BN;
    for (int i = threadIdx.x; i < N; i += blockDim.x) bias_s[i] = p.bias ? __ldg(p.bias + i) : 0.f;
  }
  __syncthreads();

  // packed < 0: ordinary tile (utterance b, rows t0 .. t0+127); packed >= 0: index of a packed tail tile
  auto tile_coords = [&](int tile, int& n0, int& b, int& t0, int& packed) {
    const int mt = tile / p.n_tiles;
    n0 = (tile - mt * p.n_tiles) * BN;
    packed = -1;
    if (p.tiles_per_utt == 0) { b = 0; t0 = mt * BM; }
    else if (mt < p.full_tiles) { b = mt / p.full; t0 = (mt - b * p.full) * BM; }
    else { packed = mt - p.full_tiles; b = packed * p.upt; t0 = p.full * BM; }
  };

  if (warp == 0) {
    if (lane == 0) {  // ---- TMA producer ----
      int n = 0;      // ring position, runs across tiles
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        int n0, b, t0, packed;
        tile_coords(tile, n0, b, t0, packed);
        for (int s = 0; s < steps; ++s, ++n) {
          const int slot = n % C::STAGES, round = n / C::STAGES;
          mbar_wait(&empty_bar[slot], (round & 1) ^ 1);
          const int j = s / kchunks, k0 = (s - j * kchunks) * C::BKE;
          uint8_t* st = tiles + (size_t)slot * C::STAGE_BYTES;
          mbar_expect_tx(&full_bar[slot], C::TX_BYTES);
          if (packed < 0) {
            tma_load_3d(st, &tmap_a, &full_bar[slot], k0, t0 + j - p.pad, b);
            if (C::SPLIT16) tma_load_3d(st + C::A_LO, &tmap_a, &full_bar[slot], k0, t0 + j - p.pad, b + p.B);
          } else {
            // eight 16-row boxes: granule g belongs to utterance b + g / gn (zero-filled past the batch or past L)
            for (int g = 0; g < 8; ++g) {
              const int u = g / p.gn, gi = g - u * p.gn;
              const bool real = u < p.upt && b + u < p.B;
              const int oob = C::SPLIT16 ? 2 * p.B : p.B;  // out of bounds in dim 2 -> the box is all zeros
              tma_load_3d(st + g * (16 * 128), &tmap_a16, &full_bar[slot], k0, t0 + gi * 16 + j - p.pad, real ? b + u : oob);
              if (C::SPLIT16) tma_load_3d(st + C::A_LO + g * (16 * 128), &tmap_a16, &full_bar[slot], k0, t0 + gi * 16 + j - p.pad, real ? b + u + p.B : oob);
            }
          }
          tma_load_3d(st + C::B_HI, &tmap_b, &full_bar[slot], k0, n0, j);
          if (PRECISE) tma_load_3d(st + C::B_LO, &tmap_b_lo, &full_bar[slot], k0, n0, j);
        }
      }
    }
  } else if (warp == 1) {
    {  // ---- MMA issuer: all 32 lanes run the loop, one lane is elected inside each tcgen05 asm ----
      int n = 0, it = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
        const int acc = it % C::NACC;
        mbar_wait(&acc_empty[acc], ((it / C::NACC) & 1) ^ 1);   // epilogue has drained this buffer (first NACC uses pass)
        tcgen05_fence_after();
        const uint32_t d = tmem_base + (uint32_t)(acc * C::ACC_STRIDE);
        for (int s = 0; s < steps; ++s, ++n) {
          const int slot = n % C::STAGES, round = n / C::STAGES;
          mbar_wait(&full_bar[slot], round & 1);
          tcgen05_fence_after();
          const uint32_t base = smem_u32(tiles + (size_t)slot * C::STAGE_BYTES);
          const uint64_t a_hi = make_sw128_kmajor_desc(base), b_hi = make_sw128_kmajor_desc(base + C::B_HI);
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {  // +32 bytes along K inside the swizzle row = +2 in descriptor units
            if (PRECISE) {
              const uint64_t a_lo = make_sw128_kmajor_desc(base + C::A_LO), b_lo = make_sw128_kmajor_desc(base + C::B_LO);
              umma_f16(d, a_lo + 2 * k, b_hi + 2 * k, C::IDESC, (s | k) != 0);  // small terms first
              umma_f16(d, a_hi + 2 * k, b_lo + 2 * k, C::IDESC, 1);
              umma_f16(d, a_hi + 2 * k, b_hi + 2 * k, C::IDESC, 1);
            } else if (HALF) {
              umma_f16(d, a_hi + 2 * k, b_hi + 2 * k, C::IDESC, (s | k) != 0);
            } else {
              umma_tf32(d, a_hi + 2 * k, b_hi + 2 * k, C::IDESC, (s | k) != 0);
            }
          }
          tcgen05_commit(&empty_bar[slot]);    // slot reusable once these MMAs have read it
        }
        tcgen05_commit(&acc_full[acc]);        // accumulator complete
      }
    }
  } else if (warp < 2 + 4 * C::GROUPS) {
    // ---- epilogue: group g = warps 2+4g .. 5+4g; warp w owns TMEM lanes 32*(w%4) .. +31;
    //      thread == output row; the groups take alternate 32-column chunks ----
    const int wq = warp & 3, grp = (warp - 2) >> 2;
    const int row = wq * 32 + lane;

    const uint32_t lane_off = (uint32_t)(wq * 32) << 16;
    const int act = p.act, ldr = p.ldr, ldo = p.ldo;
    const float* __restrict__ resid = p.resid; float* __restrict__ out = p.out;
    const bool has_res = resid != nullptr;
    float v[32];
    int it = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
      int n0, b, t0, packed;
      tile_coords(tile, n0, b, t0, packed);
      const int acc = it % C::NACC;
      int t = t0 + row;
      bool row_ok = t < p.L;                     // flat mode: L == total rows
      if (packed >= 0) {                         // packed tail tile: 16-row granule g of the tile -> utterance b + g / gn
        const int g = row >> 4, u = g / p.gn, gi = g - u * p.gn;
        b += u;
        t = t0 + gi * 16 + (row & 15);
        row_ok = u < p.upt && b < p.B && t < p.L;
      }
      const long m = (long)b * p.L + t;
      if (has_res && row_ok) {                   // residual rows -> L2 while the main loop of this tile runs
        for (int c0 = grp * 32; c0 < BN; c0 += 32 * C::GROUPS) prefetch_l2(resid + m * ldr + n0 + c0);
      }
      mbar_wait(&acc_full[acc], (it / C::NACC) & 1);
      tcgen05_fence_after();
      const uint32_t taddr = tmem_base + lane_off + (uint32_t)(acc * C::ACC_STRIDE);
      const bool to_vt = p.vt_out != nullptr && n0 >= p.vt_col0;   // tile-uniform (tile widths divide the V third)
#pragma unroll 1
      for (int c0 = grp * 32; c0 < BN; c0 += 32 * C::GROUPS) {
        // the residual is fetched first so the loads are in flight across the TMEM load; the bias comes from shared
        // memory (global loads next to their use stalled the whole epilogue: ncu long-scoreboard samples)
        float4 rv[8];
        const float4* bq = reinterpret_cast<const float4*>(bias_s + n0 + c0);
        if (to_vt) {
          __syncwarp();
          tmem_ld32(taddr + c0, v);
          if (row_ok) {
            // transposed store: for a fixed column the 32 lanes hold 32 consecutive time steps -> 128-byte rows
            const long ub = m / p.vt_L; const int ut = (int)(m - ub * p.vt_L);
            const int rel = n0 + c0 - p.vt_col0, hh = rel / p.vt_dk, d0 = rel - hh * p.vt_dk;
            float* dst = p.vt_out + ((ub * p.vt_heads + hh) * p.vt_dk + d0) * (long)p.vt_lpad + ut;
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (c0 + i < BN) dst[(long)i * p.vt_lpad] = v[i] + bias_s[n0 + c0 + i];
          }
          continue;
        }
        const bool full = c0 + 32 <= BN;          // all 32 columns of the chunk exist (always, except the N = 80 tail)
        if (has_res && row_ok) {
#pragma unroll
          for (int q = 0; q < 8; ++q)
            if (full || c0 + q * 4 < BN) rv[q] = __ldg(reinterpret_cast<const float4*>(resid + m * ldr + n0 + c0 + q * 4));
        }
        __syncwarp();
        if (!(p.debug & 1)) tmem_ld32(taddr + c0, v);
        if (row_ok && !(p.debug & 2)) {
          // flags were hoisted into registers and the activation switch sits outside the element loops: the
          // per-element predicate / constant-bank reloads of the first version made this epilogue latency-bound
#pragma unroll
          for (int q = 0; q < 8; ++q) { const float4 b4 = bq[q]; v[q * 4] += b4.x; v[q * 4 + 1] += b4.y; v[q * 4 + 2] += b4.z; v[q * 4 + 3] += b4.w; }
          if (act == ACT_RELU) {
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], 0.f);
          } else if (act == ACT_TANH) {
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = tanhf(v[i]);
          }
          if (has_res) {
#pragma unroll
            for (int q = 0; q < 8; ++q)
              if (full || c0 + q * 4 < BN) { v[q * 4] += rv[q].x; v[q * 4 + 1] += rv[q].y; v[q * 4 + 2] += rv[q].z; v[q * 4 + 3] += rv[q].w; }
          }
          if (HALF && p.out_h != nullptr) {              // fp16 copy for the next f16 GEMM: two 32-byte stores
            uint32_t h[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const __half2 t = __floats2half2_rn(fminf(fmaxf(v[2 * i], -65504.f), 65504.f), fminf(fmaxf(v[2 * i + 1], -65504.f), 65504.f));
              h[i] = *reinterpret_cast<const uint32_t*>(&t);
            }
            __half* dh = p.out_h + m * p.ldo_h + n0 + c0;
            if (full || c0 + 16 <= BN) st_global_v8_b32(dh, h);
            if (full) st_global_v8_b32(dh + 16, h + 8);
          }
          float* dst = out + m * ldo + n0 + c0;          // this thread's row: four sector-complete 32-byte stores
          if (HALF && out == nullptr) {
            // fp16-only result (the hidden activations of the conv-FFN)
          } else if (full) {
#pragma unroll
            for (int q = 0; q < 4; ++q) st_global_v8(dst + q * 8, v + q * 8);
          } else {
#pragma unroll
            for (int q = 0; q < 4; ++q)
              if (c0 + q * 8 < BN) st_global_v8(dst + q * 8, v + q * 8);
          }
        }
      }
      // this warp's TMEM reads of the tile are complete (every lane passed its last tcgen05.wait::ld)
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[acc]);
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    tmem_dealloc(tmem_base, C::TMEM_COLS);
  }
}

// 3xF16 pre-pass: x [rows][ldx] fp32 -> S[0] = hi plane, S[1] = lo plane, each [rows][K] fp16;
// hi = rn(clamp(x)), lo = rn(x - hi).  One float4 per thread and step, 8-byte stores.
__global__ void split_rows_f16_kernel(const float* __restrict__ x, int ldx, long rows, int K, __half* __restrict__ S) {
  const int kq = K >> 2;
  const long quads = rows * kq;
  __half* __restrict__ lo_plane = S + rows * K;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < quads; i += (long)gridDim.x * blockDim.x) {
    const long r = i / kq;
    const int c = (int)(i - r * kq) * 4;
    const float4 v = *reinterpret_cast<const float4*>(x + r * ldx + c);
    const float f0 = fminf(fmaxf(v.x, -65504.f), 65504.f), f1 = fminf(fmaxf(v.y, -65504.f), 65504.f);
    const float f2 = fminf(fmaxf(v.z, -65504.f), 65504.f), f3 = fminf(fmaxf(v.w, -65504.f), 65504.f);
    const __half2 h01 = __floats2half2_rn(f0, f1), h23 = __floats2half2_rn(f2, f3);
    const float2 g01 = __half22float2(h01), g23 = __half22float2(h23);
    const __half2 l01 = __floats2half2_rn(f0 - g01.x, f1 - g01.y), l23 = __floats2half2_rn(f2 - g23.x, f3 - g23.y);
    uint2 hv, lv;
    hv.x = *reinterpret_cast<const uint32_t*>(&h01); hv.y = *reinterpret_cast<const uint32_t*>(&h23);
    lv.x = *reinterpret_cast<const uint32_t*>(&l01); lv.y = *reinterpret_cast<const uint32_t*>(&l23);
    *reinterpret_cast<uint2*>(S + r * K + c) = hv;
    *reinterpret_cast<uint2*>(lo_plane + r * K + c) = lv;
  }
}

// ---- host side ----------------------------------------------------------------------------------
int sm_count() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
  }
  return n;
}

template <int BN, bool PRECISE, bool HALF = false>
int launch(const TapGemm& g, cudaStream_t st) {
  using C = Cfg<BN, PRECISE, HALF>;
  static bool configured = false;
  if (!configured) {
    FS2_CUDA_CHECK(cudaFuncSetAttribute(tap_gemm_tf32_kernel<BN, PRECISE, HALF>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM));
    configured = true;
  }
  TcParams p;
  p.K = g.K; p.taps = g.taps; p.pad = (g.taps - 1) / 2;
  p.bias = g.bias; p.resid = g.resid; p.ldr = g.ldr; p.act = g.act; p.out = g.out; p.ldo = g.ldo;
  p.out_h = HALF ? g.out_h : nullptr; p.ldo_h = g.ldo_h;
  { static int dbg = -1; if (dbg < 0) { const char* e = getenv("FS2_GEMM_DEBUG"); dbg = e ? atoi(e) : 0; } p.debug = dbg; }
  p.vt_out = g.vt_out; p.vt_col0 = g.vt_col0; p.vt_dk = g.vt_dk; p.vt_heads = g.vt_heads; p.vt_lpad = g.vt_lpad; p.vt_L = g.L;
  CUtensorMap ma, mb, mb_lo, ma16;
  int rc;
  const int esz = HALF ? 2 : 4;
  constexpr bool AH = HALF;
  constexpr int planes = HALF && PRECISE ? 2 : 1;          // 3xF16: [hi plane][lo plane], each [B*L][K] fp16 (split_rows_f16)
  const void* xa = HALF ? (PRECISE ? (const void*)g.split_ws : (const void*)g.x_h) : (const void*)g.x;
  const uint64_t row_bytes = HALF ? (PRECISE ? (uint64_t)g.K * 2 : (uint64_t)g.ldx_h * 2) : (uint64_t)g.ldx * 4;
  if (g.taps == 1) {  // flat [B*L, K]
    const uint64_t M = (uint64_t)g.B * g.L;
    p.L = (int)M; p.tiles_per_utt = 0;
    p.m_tiles = (int)((M + BM - 1) / BM);
    p.B = 1; p.full = 0; p.gn = 1; p.upt = 1; p.full_tiles = 0;
    if ((rc = make_map(&ma, xa, g.K, M, planes, row_bytes, row_bytes * M, BM, AH))) return rc;
    ma16 = ma;
  } else {            // per-utterance tiles: shifted boxes zero-fill outside [0, L)
    p.L = g.L; p.tiles_per_utt = (g.L + BM - 1) / BM;
    p.B = g.B; p.full = g.L / BM;
    int tail = g.L % BM;
    p.gn = tail ? (tail + 15) / 16 : 1;
    p.upt = tail ? 8 / p.gn : 1;
    if (tail && p.upt == 1) { p.full += 1; tail = 0; p.gn = 1; }   // tail > 64 rows: nothing to share, keep one ordinary (partly empty) tile
    p.full_tiles = p.full * g.B;
    p.m_tiles = p.full_tiles + (tail ? (g.B + p.upt - 1) / p.upt : 0);
    if ((rc = make_map(&ma, xa, g.K, g.L, (uint64_t)g.B * planes, row_bytes, row_bytes * g.L, BM, AH))) return rc;
    if ((rc = make_map(&ma16, xa, g.K, g.L, (uint64_t)g.B * planes, row_bytes, row_bytes * g.L, 16, AH))) return rc;
  }
  p.n_tiles = g.N / BN;
  const void* w_hi = PRECISE ? (const void*)g.w_hi_h : HALF ? (const void*)g.w_h : (const void*)g.w;
  if ((rc = make_map(&mb, w_hi, g.K, g.N, g.taps, (uint64_t)g.K * esz, (uint64_t)g.K * esz * g.N, BN, HALF))) return rc;
  if ((rc = make_map(&mb_lo, PRECISE ? (const void*)g.w_lo_h : w_hi, g.K, g.N, g.taps, (uint64_t)g.K * esz, (uint64_t)g.K * esz * g.N, BN, HALF))) return rc;
  const int total = p.m_tiles * p.n_tiles;
  const int grid = total < sm_count() ? total : sm_count();
  tap_gemm_tf32_kernel<BN, PRECISE, HALF><<<grid, C::THREADS, C::SMEM, st>>>(ma, mb, mb_lo, ma16, p);
  FS2_LAUNCH_CHECK();
  return FS2_OK;
}

int check_output(const TapGemm& g, const char* who) {
  FS2_REQUIRE((g.taps & 1) == 1, "%s: taps must be odd", who);
  FS2_REQUIRE(g.N % 16 == 0 && g.N <= 2048, "%s: N (%d) must be a multiple of 16 and fit the staged bias vector (2048)", who, g.N);
  FS2_REQUIRE(!g.resid || g.ldr % 4 == 0, "%s: row strides must be 16-byte multiples", who);
  FS2_REQUIRE(g.ldo % 8 == 0 && (reinterpret_cast<uintptr_t>(g.out) & 31) == 0, "%s: output rows must be 32-byte aligned (256-bit stores)", who);
  return FS2_OK;
}
int check_common(const TapGemm& g, const char* who) {
  FS2_REQUIRE(g.K % 4 == 0, "%s: K (%d) must be a multiple of 4", who, g.K);
  FS2_REQUIRE(g.ldx % 4 == 0, "%s: row strides must be 16-byte multiples", who);
  FS2_REQUIRE((reinterpret_cast<uintptr_t>(g.x) & 15) == 0 && (reinterpret_cast<uintptr_t>(g.w) & 15) == 0 && g.out != nullptr,
              "%s: operands must be 16-byte aligned", who);
  return check_output(g, who);
}

}  // namespace

int tap_gemm_tf32(const TapGemm& g, cudaStream_t st) {
  int rc = check_common(g, "tap_gemm_tf32");
  if (rc) return rc;
  if ((long)g.B * g.L == 0) return FS2_OK;
  {
    static int force = -1;   // FS2_GEMM_BN: tile-width override for experiments
    if (force < 0) { const char* e = getenv("FS2_GEMM_BN"); force = e ? atoi(e) : 0; }
    if (force == 128 && g.N % 128 == 0) return launch<128, false>(g, st);
    if (force == 64 && g.N % 64 == 0) return launch<64, false>(g, st);
  }
  if (g.N % 256 == 0) return launch<256, false>(g, st);
  if (g.N % 192 == 0) return launch<192, false>(g, st);
  if (g.N % 128 == 0) return launch<128, false>(g, st);
  if (g.N % 96 == 0) return launch<96, false>(g, st);
  if (g.N % 80 == 0) return launch<80, false>(g, st);
  if (g.N % 64 == 0) return launch<64, false>(g, st);
  set_error("tap_gemm_tf32: N=%d has no supported tile width", g.N);
  return FS2_ERR_INVALID;
}

int tap_gemm_3xtf32(const TapGemm& g, cudaStream_t st) {
  int rc = check_common(g, "tap_gemm_3xtf32");
  if (rc) return rc;
  if ((long)g.B * g.L == 0) return FS2_OK;
  FS2_REQUIRE(g.w_hi_h && g.w_lo_h && g.split_ws, "tap_gemm_3xtf32: split weights / activation scratch missing");
  FS2_REQUIRE(g.K % 8 == 0, "tap_gemm_3xtf32: K (%d) must be a multiple of 8", g.K);
  if (!g.split_ready) {
    // pre-pass: fp32 activations -> fp16 hi / lo planes (rows of K contiguous halfs)
    const long rows = (long)g.B * g.L;
    const long quads = rows * (g.K / 4);
    long blocks = (quads + 255) / 256;
    split_rows_f16_kernel<<<(int)(blocks > 148 * 16 ? 148 * 16 : blocks), 256, 0, st>>>(g.x, g.ldx, rows, g.K, g.split_ws);
    FS2_LAUNCH_CHECK();
  }
  // narrower tiles than the plain kernel: the stage holds four operand tiles and the encoder's M is small
  if (g.N % 128 == 0) return launch<128, true, true>(g, st);
  if (g.N % 96 == 0) return launch<96, true, true>(g, st);
  if (g.N % 80 == 0) return launch<80, true, true>(g, st);
  if (g.N % 64 == 0) return launch<64, true, true>(g, st);
  set_error("tap_gemm_3xtf32: N=%d has no supported tile width", g.N);
  return FS2_ERR_INVALID;
}

int tap_gemm_f16(const TapGemm& g, cudaStream_t st) {
  const char* who = "tap_gemm_f16";
  FS2_REQUIRE(g.x_h && g.w_h && (g.out || g.out_h), "%s: fp16 operands / an output missing", who);
  FS2_REQUIRE(g.K % 8 == 0 && g.ldx_h % 8 == 0, "%s: K (%d) and the fp16 row stride must be multiples of 8", who, g.K);
  FS2_REQUIRE((reinterpret_cast<uintptr_t>(g.x_h) & 15) == 0 && (reinterpret_cast<uintptr_t>(g.w_h) & 15) == 0, "%s: operands must be 16-byte aligned", who);
  FS2_REQUIRE(!g.out_h || (g.ldo_h % 16 == 0 && (reinterpret_cast<uintptr_t>(g.out_h) & 31) == 0), "%s: fp16 output rows must be 32-byte aligned", who);
  FS2_REQUIRE(!g.ln_gamma, "%s: the LayerNorm epilogue lives in gemm_ln_tc.cu", who);
  int rc = check_output(g, who);
  if (rc) return rc;
  if ((long)g.B * g.L == 0) return FS2_OK;
  if (g.N % 256 == 0) return launch<256, false, true>(g, st);
  if (g.N % 192 == 0) return launch<192, false, true>(g, st);
  if (g.N % 128 == 0) return launch<128, false, true>(g, st);
  if (g.N % 80 == 0) return launch<80, false, true>(g, st);
  set_error("%s: N=%d has no supported tile width (multiples of 80, 128 or 192)", who, g.N);
  return FS2_ERR_INVALID;
}

namespace {
__global__ void split_f16_kernel(const float* __restrict__ src, __half* __restrict__ hi, __half* __restrict__ lo, long n) {
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    const float x = fminf(fmaxf(src[i], -65504.f), 65504.f);
    const __half h = __float2half_rn(x);
    hi[i] = h; lo[i] = __float2half_rn(x - __half2float(h));
  }
}
}  // namespace

int split_f16(const float* src, __half* hi, __half* lo, long n, cudaStream_t st) {
  if (n == 0) return FS2_OK;
  long blocks = (n + 255) / 256;
  split_f16_kernel<<<(int)(blocks > 1184 ? 1184 : blocks), 256, 0, st>>>(src, hi, lo, n);
  FS2_LAUNCH_CHECK();
  return FS2_OK;
}

}  // namespace fs2
