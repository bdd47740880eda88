// Tensor-core "tap GEMM" for sm_100a: tcgen05.mma (kind::tf32 or kind::f16) with TMEM accumulators,
// operands staged in shared memory by TMA (cp.async.bulk.tensor, 128-byte swizzle), mbarrier
// pipelines, persistent CTAs.
//
//   out[b,t,n] = act( sum_{j<taps} sum_{k<K} x[b, t+j-pad, k] * w[j][n][k] + bias[n] ) (+ resid[b,t,n])
//
// Same contract as gemm_fp32.cu (the CUDA-core family).  Three instantiation families:
//   <BN, false, false>  plain tf32 (one MMA per product) on the fp32 activations themselves: the decoder side in
//                       FS2_MATH_TF32;
//   <BN, false, true>   kind::f16 on the hi planes of activations and weights (decoder side in FS2_MATH_F16);
//                       see "HALF" below;
//   <BN, true,  true>   error-compensated "3xF16" on hi + lo planes (encoder and predictors in every tensor-core mode,
//                       everything in FS2_MATH_3XTF32): fp32-class results; their outputs feed round() / bucketize(),
//                       where 10-bit-mantissa noise (~1e-3) would flip integers; see "PRECISE && HALF" below.
// Operand planes (common.cuh): activations are pre-scaled by kPlaneScale, weights by a per-layer power of two; the
// epilogue multiplies the accumulator by the exact inverse (oscale) in the same FMA that adds the bias.  Results leave as
// fp32 rows and / or as the operand planes of the next contraction (no separate split / conversion pass).
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
//   smem ring   : warp 0 (TMA producer)  <-> warp 1 (MMA issuer), full/empty mbarriers; both run warp-convergent loops with one
//                 lane elected inside each asm, and compute a stage's operands before waiting for it
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
// PRECISE && HALF ("3xF16", the error-compensated family): the producer of x writes it as two fp16 planes, hi = rn(s x)
// and lo = rn(s x - hi), the weights are split the same way at load time (with their own power-of-two scale, so the lo
// plane of a trained layer's small weights is a normal fp16, not a subnormal), and the GEMM loads all four operand tiles
// by TMA.  The three products run as TWO instructions per K = 16 step: the B stage holds [b_hi | b_lo] as one 2*BN-row
// tile, so  a_hi . [b_hi | b_lo]  is a single N = 2*BN MMA into accumulator columns [0,BN) (main term) and [BN,2BN)
// (correction), and  a_lo . b_hi  (N = BN) accumulates into the correction columns; the epilogue adds the two.  Versus
// three N = BN MMAs into one accumulator: 20 % less shared-memory operand traffic (the N = BN 3xF16 loop sat at the
// 128 B/clk port limit), a third of the issue slots, and the main sum sees K/16 round-toward-zero accumulation steps
// instead of 3K/16.  (BN = 80 keeps three MMAs into one accumulator: its correction block would not start on a
// 32-column TMEM boundary.)  The lo plane is addressed through the same tensor map: plane stride = B*L rows, i.e.
// utterance index b + B.
//
// HALF = true, PRECISE = false (FS2_MATH_F16, the decoder side): the same pipeline on the hi planes of the activations and weights
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
  float a_inv; const float* w_inv;   // accumulator scale = a_inv * (w_inv ? *w_inv : 1): undoes the operand planes' pre-scaling
  __half* outp; __half* outp_lo; int ldo_p;   // result as operand planes (hi; lo when the consumer is 3xF16); out may then be null
  // optional: columns >= vt_col0 are the V third of a q|k|v projection and are stored transposed,
  // vt[(b*heads + h)*dk + d][t] with row pitch vt_lpad, for the attention kernel's K-major P.V operand
  float* vt_out; __half* vtp; __half* vtp_lo; int vt_col0, vt_dk, vt_heads, vt_lpad, vt_L;
  int debug;   // FS2_GEMM_DEBUG: bit0 skip tcgen05.ld, bit1 skip stores, bit2 V third untransposed (profiling experiments only)
  int prefetch;  // FS2_GEMM_PREFETCH (default 0: measured slower, see the producer loop): L2 prefetch of the next tile's activation rows
};

constexpr int pow2_at_least(int x) { return x <= 32 ? 32 : x <= 64 ? 64 : x <= 128 ? 128 : x <= 256 ? 256 : 512; }

template <int BN, bool PRECISE, bool HALF = false>
struct Cfg {
  static constexpr int B_BYTES = BN * BK * 4;
  static constexpr int STAGE_BYTES = (PRECISE ? 2 : 1) * (A_BYTES + B_BYTES);   // [A(hi)][A lo][B hi][B lo]
  static constexpr bool SPLIT16 = PRECISE;                // A hi / lo are fp16 planes in global memory, loaded like B hi / lo
  static_assert(!PRECISE || HALF, "the error-compensated family is 3xF16");
  static constexpr int STAGES = (RING_BUDGET / STAGE_BYTES) > 8 ? 8 : (RING_BUDGET / STAGE_BYTES);
  static constexpr bool WIDE = PRECISE && (BN % 32 == 0);  // main + correction accumulators, a_hi . [b_hi | b_lo] as one MMA
  static constexpr int ACC_STRIDE = pow2_at_least(WIDE ? 2 * BN : BN);     // TMEM columns per accumulator buffer
  static constexpr int NACC = 512 / ACC_STRIDE > 4 ? 4 : 512 / ACC_STRIDE;   // accumulator buffers in flight (2 for BN > 128, else 4)
  static constexpr int TMEM_COLS = NACC * ACC_STRIDE;
  static constexpr int GROUPS = 2;       // epilogue warp groups (4 warps each), alternate 32-column chunks
  static constexpr int THREADS = 64 + GROUPS * 128;
  static constexpr size_t SMEM = (size_t)STAGES * STAGE_BYTES + STAGING_BYTES + 1024 + 512;
  static constexpr uint32_t IDESC = HALF ? idesc_f16(BM, BN) : idesc_tf32(BM, BN);
  static constexpr uint32_t IDESC_WIDE = idesc_f16(BM, WIDE ? 2 * BN : BN);
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

  const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;
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
  pdl_wait();                                              // everything above overlapped the previous kernel's tail
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
    // ---- TMA producer: the whole warp runs the loop, one lane is elected inside each asm.  Everything a stage's loads need
    // (coordinates, shared-memory and barrier addresses) is computed BEFORE the wait for its slot, and the tap / K-chunk indices
    // are carried as counters: what used to sit between "slot free" and "loads issued" (an integer division, address arithmetic,
    // an ELECT / vote loop per UTMALDG: ~100 of the loop's 140 instructions) is part of the refill latency of a 3-stage ring.
    const uint32_t tiles_addr = smem_u32(tiles), full_addr = smem_u32(full_bar);
    int n = 0;      // ring position, runs across tiles
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      int n0, b, t0, packed;
      tile_coords(tile, n0, b, t0, packed);
      // Experiment (FS2_GEMM_PREFETCH=1, off by default): L2 prefetch of the NEXT tile's activation rows, one K chunk per
      // pipeline step, by the one CTA whose next tile is the first column tile of its row tile.  Hypothesis: first-touch
      // activations arrive with DRAM latency that the 3 - 8 stage ring cannot cover (ncu of the K = 384 q|k|v projection: TMA,
      // MMA and epilogue warps each wait ~30 %, nothing saturated).  Measured on c2 / 3xF16: q|k|v 0.485 -> 0.51 ms, w_2
      // (cluster kernel, same idea) 0.53 -> 0.64 ms: the extra TMA traffic costs more than the latency it hides.
      const int nxt = tile + (int)gridDim.x;
      int pn0 = 0, pb = 0, pt0 = 0, ppacked = 0;
      bool pf = p.prefetch && nxt < total_tiles && (nxt % p.n_tiles) == 0;
      if (pf) { tile_coords(nxt, pn0, pb, pt0, ppacked); pf = ppacked < 0; }
      int j = 0, kc = 0;                        // tap, K chunk of step s
      for (int s = 0; s < steps; ++s, ++n) {
        if (pf && s < 2 * kchunks && (s < kchunks || p.taps > 1) && lane == 0) {
          // rows [t0 - pad, +128) in the first pass; a k > 1 convolution also reads up to row t0 + 127 + pad: second pass
          const int pass = s < kchunks ? 0 : 1, pkc = s - pass * kchunks;
          const int r = pt0 + (pass ? p.pad : -p.pad);
          tma_prefetch_3d(&tmap_a, pkc * C::BKE, r, pb);
          if (C::SPLIT16) tma_prefetch_3d(&tmap_a, pkc * C::BKE, r, pb + p.B);
        }
        const int slot = n % C::STAGES, round = n / C::STAGES;
        const int k0 = kc * C::BKE, row = t0 + j - p.pad;
        const uint32_t st = tiles_addr + (uint32_t)slot * C::STAGE_BYTES, fb = full_addr + (uint32_t)slot * 8u;
        pin_before(st, fb, k0, row);
        mbar_wait(&empty_bar[slot], (round & 1) ^ 1);
        mbar_expect_tx_elect(fb, C::TX_BYTES);
        if (packed < 0) {
          tma_load_3d_elect(st, &tmap_a, fb, k0, row, b);
          if (C::SPLIT16) tma_load_3d_elect(st + C::A_LO, &tmap_a, fb, k0, row, b + p.B);
        } else {
          // eight 16-row boxes: granule g belongs to utterance b + g / gn (zero-filled past the batch or past L)
          for (int g = 0; g < 8; ++g) {
            const int u = g / p.gn, gi = g - u * p.gn;
            const bool real = u < p.upt && b + u < p.B;
            const int oob = C::SPLIT16 ? 2 * p.B : p.B;  // out of bounds in dim 2 -> the box is all zeros
            tma_load_3d_elect(st + g * (16 * 128), &tmap_a16, fb, k0, row + gi * 16, real ? b + u : oob);
            if (C::SPLIT16) tma_load_3d_elect(st + C::A_LO + g * (16 * 128), &tmap_a16, fb, k0, row + gi * 16, real ? b + u + p.B : oob);
          }
        }
        tma_load_3d_elect(st + C::B_HI, &tmap_b, fb, k0, n0, j);
        if (PRECISE) tma_load_3d_elect(st + C::B_LO, &tmap_b_lo, fb, k0, n0, j);
        if (++kc == kchunks) { kc = 0; ++j; }
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
          // descriptors of the stage before the wait for its data (they are part of the data-landed -> first-MMA latency otherwise)
          const uint32_t base = smem_u32(tiles + (size_t)slot * C::STAGE_BYTES);
          const uint64_t a_hi = make_sw128_kmajor_desc(base), b_hi = make_sw128_kmajor_desc(base + C::B_HI);
          const uint64_t a_lo_w = make_sw128_kmajor_desc(base + C::A_LO), b_lo_w = make_sw128_kmajor_desc(base + C::B_LO);
          pin_before64(a_hi, b_hi, a_lo_w, b_lo_w);
          mbar_wait(&full_bar[slot], round & 1);
          tcgen05_fence_after();
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {  // +32 bytes along K inside the swizzle row = +2 in descriptor units
            if (C::WIDE) {
              const uint64_t a_lo = a_lo_w;
              umma_f16(d, a_hi + 2 * k, b_hi + 2 * k, C::IDESC_WIDE, (s | k) != 0);   // [0,BN) += a_hi b_hi, [BN,2BN) += a_hi b_lo
              umma_f16(d + BN, a_lo + 2 * k, b_hi + 2 * k, C::IDESC, 1);              // [BN,2BN) += a_lo b_hi
            } else if (PRECISE) {
              const uint64_t a_lo = a_lo_w, b_lo = b_lo_w;
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
    const float oscale = p.a_inv * (p.w_inv ? __ldg(p.w_inv) : 1.0f);   // exact power of two (1 in the tf32 family)
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
      const bool to_vt = (p.vt_out != nullptr || p.vtp != nullptr) && n0 >= p.vt_col0;   // tile-uniform (tile widths divide the V third)
#pragma unroll 1
      for (int c0 = grp * 32; c0 < BN; c0 += 32 * C::GROUPS) {
        // the residual is fetched first so the loads are in flight across the TMEM load; the bias comes from shared
        // memory (global loads next to their use stalled the whole epilogue: ncu long-scoreboard samples)
        float4 rv[8];
        const float4* bq = reinterpret_cast<const float4*>(bias_s + n0 + c0);
        const bool full = c0 + 32 <= BN;          // all 32 columns of the chunk exist (always, except the N = 80 tail)
        if (has_res && row_ok && !to_vt) {
#pragma unroll
          for (int q = 0; q < 8; ++q)
            if (full || c0 + q * 4 < BN) rv[q] = __ldg(reinterpret_cast<const float4*>(resid + m * ldr + n0 + c0 + q * 4));
        }
        __syncwarp();
        if (!(p.debug & 1)) {
          if (C::WIDE) {                          // main + correction accumulators
            float v2[32];
            tmem_ld32_nowait(taddr + c0, v); tmem_ld32_nowait(taddr + BN + c0, v2);
            tmem_ld_wait_pin<32>(v); tmem_ld_wait_pin<32>(v2);
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] += v2[i];
          } else {
            tmem_ld32(taddr + c0, v);
          }
        }
        if (!row_ok || (p.debug & 2)) continue;
        // flags were hoisted into registers and the activation switch sits outside the element loops: the
        // per-element predicate / constant-bank reloads of the first version made this epilogue latency-bound
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float4 b4 = bq[q];
          v[q * 4] = fmaf(v[q * 4], oscale, b4.x); v[q * 4 + 1] = fmaf(v[q * 4 + 1], oscale, b4.y);
          v[q * 4 + 2] = fmaf(v[q * 4 + 2], oscale, b4.z); v[q * 4 + 3] = fmaf(v[q * 4 + 3], oscale, b4.w);
        }
        if (to_vt && !(p.debug & 4)) {
          // transposed store: for a fixed column the 32 lanes hold 32 consecutive time steps -> contiguous 128-byte
          // (fp32) / 64-byte (fp16 plane) runs
          const long ub = m / p.vt_L; const int ut = (int)(m - ub * p.vt_L);
          const int rel = n0 + c0 - p.vt_col0, hh = rel / p.vt_dk, d0 = rel - hh * p.vt_dk;
          const long o = ((ub * p.vt_heads + hh) * p.vt_dk + d0) * (long)p.vt_lpad + ut;
          if (p.vt_out != nullptr) {
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (c0 + i < BN) p.vt_out[o + (long)i * p.vt_lpad] = v[i];
          } else {
#pragma unroll
            for (int i = 0; i < 32; i += 2) {
              if (c0 + i >= BN) break;
              uint32_t hi, lo;
              split_pair(v[i], v[i + 1], hi, lo);
              const __half2 h2 = *reinterpret_cast<const __half2*>(&hi), l2 = *reinterpret_cast<const __half2*>(&lo);
              p.vtp[o + (long)i * p.vt_lpad] = __low2half(h2); p.vtp[o + (long)(i + 1) * p.vt_lpad] = __high2half(h2);
              if (p.vtp_lo != nullptr) { p.vtp_lo[o + (long)i * p.vt_lpad] = __low2half(l2); p.vtp_lo[o + (long)(i + 1) * p.vt_lpad] = __high2half(l2); }
            }
          }
          continue;
        }
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
        if (HALF && p.outp != nullptr) {              // operand planes of the next contraction: two 32-byte stores per plane
          uint32_t h[16], l[16];
          if (p.outp_lo != nullptr) {
#pragma unroll
            for (int i = 0; i < 16; ++i) split_pair(v[2 * i], v[2 * i + 1], h[i], l[i]);
          } else {
#pragma unroll
            for (int i = 0; i < 16; ++i) h[i] = hi_pair(v[2 * i], v[2 * i + 1]);
          }
          long off = m * p.ldo_p + n0 + c0;
          __half* ob = p.outp; __half* ob_lo = p.outp_lo;
          if (to_vt) {   // FS2_GEMM_DEBUG bit 2 (timing experiment only): the V third stored UNtransposed into the V^T buffer (wrong layout)
            ob = p.vtp; ob_lo = p.vtp_lo; off = m * (long)(p.vt_dk * p.vt_heads) + n0 - p.vt_col0 + c0;
          }
          if (full || c0 + 16 <= BN) st_global_v8_b32(ob + off, h);
          if (full) st_global_v8_b32(ob + off + 16, h + 8);
          if (ob_lo != nullptr) {
            if (full || c0 + 16 <= BN) st_global_v8_b32(ob_lo + off, l);
            if (full) st_global_v8_b32(ob_lo + off + 16, l + 8);
          }
        }
        if (out != nullptr) {
          float* dst = out + m * ldo + n0 + c0;          // this thread's row: four sector-complete 32-byte stores
          if (full) {
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
  pdl_trigger();                               // this CTA's work is done: the next kernel may start launching (FS2_PDL)
  if (warp == 1) {
    tcgen05_fence_after();
    tmem_dealloc(tmem_base, C::TMEM_COLS);
  }
}

// Pre-pass for tensors that enter the library as fp32 (LengthRegulator output, single-operator test entries):
// x [rows][ldx] fp32 -> S[0] = hi plane, S[1] = lo plane, each [rows][K] fp16, scaled by kPlaneScale.
// One float4 per thread and step, 8-byte stores.
__global__ void split_rows_f16_kernel(const float* __restrict__ x, int ldx, long rows, int K, __half* __restrict__ S) {
  pdl_trigger(); pdl_wait();
  const int kq = K >> 2;
  const long quads = rows * kq;
  __half* __restrict__ lo_plane = S + rows * K;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < quads; i += (long)gridDim.x * blockDim.x) {
    const long r = i / kq;
    const int c = (int)(i - r * kq) * 4;
    const float4 v = *reinterpret_cast<const float4*>(x + r * ldx + c);
    uint2 hv, lv;
    split_pair(v.x, v.y, hv.x, lv.x);
    split_pair(v.z, v.w, hv.y, lv.y);
    *reinterpret_cast<uint2*>(S + r * K + c) = hv;
    *reinterpret_cast<uint2*>(lo_plane + r * K + c) = lv;
  }
}

// ---- host side ----------------------------------------------------------------------------------
template <int BN, bool PRECISE, bool HALF = false>
int launch(const TapGemm& g, cudaStream_t st) {
  using C = Cfg<BN, PRECISE, HALF>;
  static unsigned long long configured = 0;   // per-device bit mask
  int rc;
  if ((rc = ensure_smem_attr(tap_gemm_tf32_kernel<BN, PRECISE, HALF>, C::SMEM, &configured))) return rc;
  TcParams p;
  p.K = g.K; p.taps = g.taps; p.pad = (g.taps - 1) / 2;
  p.bias = g.bias; p.resid = g.resid; p.ldr = g.ldr; p.act = g.act; p.out = g.out; p.ldo = g.ldo;
  const long rows = (long)g.B * g.L;
  p.a_inv = HALF ? g.a_inv : 1.0f; p.w_inv = HALF ? g.w_inv : nullptr;
  p.outp = HALF ? g.outp : nullptr; p.ldo_p = g.ldo_p;
  p.outp_lo = (HALF && g.outp && g.outp_lo) ? g.outp + rows * g.ldo_p : nullptr;
  { static int dbg = -1; if (dbg < 0) { const char* e = getenv("FS2_GEMM_DEBUG"); dbg = e ? atoi(e) : 0; } p.debug = dbg; }
  { static int pfe = -1; if (pfe < 0) { const char* e = getenv("FS2_GEMM_PREFETCH"); pfe = e ? atoi(e) : 0; } p.prefetch = pfe; }
  p.vt_out = HALF ? nullptr : g.vt_out; p.vtp = HALF ? g.vtp : nullptr;
  p.vtp_lo = (HALF && g.vtp && g.outp_lo) ? g.vtp + (long)g.B * g.vt_heads * g.vt_dk * g.vt_lpad : nullptr;
  p.vt_col0 = g.vt_col0; p.vt_dk = g.vt_dk; p.vt_heads = g.vt_heads; p.vt_lpad = g.vt_lpad; p.vt_L = g.L;
  CUtensorMap ma, mb, mb_lo, ma16;
  const int esz = HALF ? 2 : 4;
  constexpr bool AH = HALF;
  constexpr int planes = HALF && PRECISE ? 2 : 1;          // 3xF16: [hi plane][lo plane], each [B*L][K] fp16
  const void* xa = HALF ? (const void*)g.xp : (const void*)g.x;
  const uint64_t row_bytes = HALF ? (uint64_t)g.K * 2 : (uint64_t)g.ldx * 4;
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
  const void* w_hi = HALF ? (const void*)g.w_hi : (const void*)g.w;
  if ((rc = make_map(&mb, w_hi, g.K, g.N, g.taps, (uint64_t)g.K * esz, (uint64_t)g.K * esz * g.N, BN, HALF))) return rc;
  if ((rc = make_map(&mb_lo, PRECISE ? (const void*)g.w_lo : w_hi, g.K, g.N, g.taps, (uint64_t)g.K * esz, (uint64_t)g.K * esz * g.N, BN, HALF))) return rc;
  const int total = p.m_tiles * p.n_tiles;
  const int grid = total < sm_count_current() ? total : sm_count_current();
  FS2_CUDA_CHECK(launch_pdl(tap_gemm_tf32_kernel<BN, PRECISE, HALF>, dim3(grid), dim3(C::THREADS), C::SMEM, st, ma, mb, mb_lo, ma16, p));
  FS2_LAUNCH_CHECK();
  return FS2_OK;
}

int check_output(const TapGemm& g, const char* who) {
  FS2_REQUIRE((g.taps & 1) == 1, "%s: taps must be odd", who);
  FS2_REQUIRE(g.N % 16 == 0 && g.N <= 2048, "%s: N (%d) must be a multiple of 16 and fit the staged bias vector (2048)", who, g.N);
  FS2_REQUIRE(!g.resid || g.ldr % 4 == 0, "%s: row strides must be 16-byte multiples", who);
  FS2_REQUIRE(!g.out || (g.ldo % 8 == 0 && (reinterpret_cast<uintptr_t>(g.out) & 31) == 0), "%s: output rows must be 32-byte aligned (256-bit stores)", who);
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

// kind::f16 on the hi planes (g.precise == false) or 3xF16 on hi + lo planes
int tap_gemm_planes(const TapGemm& g, cudaStream_t st) {
  const char* who = g.precise ? "tap_gemm_planes(3xF16)" : "tap_gemm_planes(f16)";
  FS2_REQUIRE(g.xp && g.w_hi && (!g.precise || g.w_lo) && (g.out || g.outp || g.vtp), "%s: operand planes / an output missing", who);
  FS2_REQUIRE(g.K % 8 == 0, "%s: K (%d) must be a multiple of 8 (16-byte plane rows)", who, g.K);
  FS2_REQUIRE((reinterpret_cast<uintptr_t>(g.xp) & 15) == 0 && (reinterpret_cast<uintptr_t>(g.w_hi) & 15) == 0, "%s: operands must be 16-byte aligned", who);
  FS2_REQUIRE(!g.outp || (g.ldo_p % 16 == 0 && (reinterpret_cast<uintptr_t>(g.outp) & 31) == 0 && (((long)g.B * g.L * g.ldo_p) % 16) == 0),
              "%s: output plane rows must be 32-byte aligned", who);
  {
    const int bn = g.precise ? 128 : (g.N % 256 == 0 ? 256 : g.N % 192 == 0 ? 192 : 128);
    FS2_REQUIRE(!g.vtp || (g.vt_lpad % 8 == 0 && g.N % bn == 0 && g.vt_col0 % bn == 0 && g.vt_dk % 32 == 0),
                "%s: transposed V planes need a 16-byte row pitch and tile-aligned thirds", who);
  }
  FS2_REQUIRE(!g.ln_gamma, "%s: the LayerNorm epilogue lives in gemm_ln_tc.cu", who);
  int rc = check_output(g, who);
  if (rc) return rc;
  if ((long)g.B * g.L == 0) return FS2_OK;
  if (g.precise) {
    // narrower tiles than the plain kernel: the stage holds four operand tiles and the encoder's M is small
    if (g.N % 128 == 0) return launch<128, true, true>(g, st);
    if (g.N % 96 == 0) return launch<96, true, true>(g, st);
    if (g.N % 80 == 0) return launch<80, true, true>(g, st);
    if (g.N % 64 == 0) return launch<64, true, true>(g, st);
  } else {
    if (g.N % 256 == 0) return launch<256, false, true>(g, st);
    if (g.N % 192 == 0) return launch<192, false, true>(g, st);
    if (g.N % 128 == 0) return launch<128, false, true>(g, st);
    if (g.N % 80 == 0) return launch<80, false, true>(g, st);
  }
  set_error("%s: N=%d has no supported tile width", who, g.N);
  return FS2_ERR_INVALID;
}

int split_rows(const float* x, int ldx, long rows, int K, __half* planes, cudaStream_t st) {
  if (rows == 0) return FS2_OK;
  FS2_REQUIRE(K % 4 == 0 && ldx % 4 == 0, "split_rows: K and the row stride must be multiples of 4");
  const long quads = rows * (K / 4);
  long blocks = (quads + 255) / 256;
  FS2_CUDA_CHECK(launch_pdl(split_rows_f16_kernel, dim3((unsigned)(blocks > 148 * 16 ? 148 * 16 : blocks)), dim3(256), 0, st, x, ldx, rows, K, planes));
  FS2_LAUNCH_CHECK();
  return FS2_OK;
}

namespace {
// weights: hi = rn(s w), lo = rn(s w - hi) with the layer's power-of-two scale s (device scalar; null = 1)
__global__ void split_f16_kernel(const float* __restrict__ src, __half* __restrict__ hi, __half* __restrict__ lo, long n,
                                 const float* __restrict__ scale) {
  const float s = scale ? __ldg(scale) : 1.0f;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    const float x = fminf(fmaxf(src[i] * s, -65504.f), 65504.f);
    const __half h = __float2half_rn(x);
    hi[i] = h; lo[i] = __float2half_rn(x - __half2float(h));
  }
}
// one CTA: max |w| -> s = 2^k with s * max in [2^13, 2^14)  (fp16 max is 2^16: headroom for rounding, none of the
// weight's lo plane below ~2^-17 of the layer maximum is subnormal)
__global__ void weight_scale_kernel(const float* __restrict__ w, long n, float* __restrict__ scale, float* __restrict__ inv) {
  __shared__ float red[32];
  float m = 0.f;
  for (long i = threadIdx.x; i < n; i += blockDim.x) m = fmaxf(m, fabsf(w[i]));
  m = warp_max(m);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x < 32) {
    m = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
    m = warp_max(m);
    if (threadIdx.x == 0) {
      int k = 0;
      if (m > 0.f && isfinite(m)) {
        int e;
        frexpf(m, &e);          // m = f * 2^e, f in [0.5, 1)  ->  m * 2^(14 - e) in [2^13, 2^14)
        k = 14 - e;
        k = k > 60 ? 60 : (k < -60 ? -60 : k);
      }
      scale[0] = ldexpf(1.0f, k); inv[0] = ldexpf(1.0f, -k);
    }
  }
}
}  // namespace

namespace {
__global__ void planes_to_rows_kernel(const __half* __restrict__ planes, long n, float* __restrict__ out) {
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x)
    out[i] = (__half2float(planes[i]) + __half2float(planes[n + i])) * kPlaneInv;
}
}  // namespace

// test helper: planes [2][n] (hi, lo; scaled by kPlaneScale) -> fp32
int planes_to_rows(const __half* planes, long n, float* out, cudaStream_t st) {
  if (n == 0) return FS2_OK;
  long blocks = (n + 255) / 256;
  planes_to_rows_kernel<<<(int)(blocks > 1184 ? 1184 : blocks), 256, 0, st>>>(planes, n, out);
  FS2_LAUNCH_CHECK();
  return FS2_OK;
}

int split_f16(const float* src, __half* hi, __half* lo, long n, const float* scale, cudaStream_t st) {
  if (n == 0) return FS2_OK;
  long blocks = (n + 255) / 256;
  split_f16_kernel<<<(int)(blocks > 1184 ? 1184 : blocks), 256, 0, st>>>(src, hi, lo, n, scale);
  FS2_LAUNCH_CHECK();
  return FS2_OK;
}

int weight_scale(const float* w, long n, float* scale, float* inv, cudaStream_t st) {
  weight_scale_kernel<<<1, 1024, 0, st>>>(w, n, scale, inv);
  FS2_LAUNCH_CHECK();
  return FS2_OK;
}

}  // namespace fs2
