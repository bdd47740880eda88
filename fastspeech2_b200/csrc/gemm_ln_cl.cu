// Row-complete GEMM + residual + LayerNorm on operand planes, as a 2-CTA cluster that splits the row:
//
//   out[m,:] = LayerNorm_C( oscale * (xp[m,:] . W^T) + bias + resid[m,:] ) * gamma + beta          (C = 256 or 384)
//
// `x = norm(x + linear_out(ctx))` and `x = norm(x + w_2(hid))` of EncoderLayer.forward (core/encoder.py:60-69) for the
// plane families: kind::f16 on the hi planes (X3 = false) or error-compensated 3xF16 on hi + lo planes (X3 = true, the
// fp32-class mode).  The result leaves as fp32 rows (the next residual) and as the operand planes of the contraction
// that consumes it next, so neither the projection's output nor the LayerNorm's input ever exists in HBM.
//
// Why a cluster: LayerNorm needs complete rows, and a 128 x 384 fp32 accumulator takes 384 of the 512 TMEM columns, so a
// single CTA cannot double-buffer it -- main loop and epilogue run back to back and the epilogue has to keep y in TMEM
// across three passes (gemm_ln_tc.cu, the kind::tf32 variant).  Here CTA r of a pair owns columns [r C/2, (r+1) C/2) of
// the same 128 rows:
//   * two C/2-column accumulators per CTA (2 x 192 <= 512): the tensor core runs tile i+1 while tile i is normalised;
//   * a thread owns half a row of its CTA's half (96 or 64 values): y stays in REGISTERS, TMEM is read once and released
//     at once, no tcgen05.st;
//   * row statistics: every thread reduces its values to (mean, M2) and the four owners of a row (two per CTA) merge them
//     with Chan's parallel-variance formula -- one exchange per tile, written into both CTAs' shared memory (local store +
//     st.shared::cluster) and published with a cluster-scope mbarrier arrive (release) / wait (acquire);
//   * the residual arrives by TMA (128-byte-swizzled 128 x 32 fp32 boxes, a ring refilled by its own producer warp while
//     the main loop runs) instead of row-strided global loads that sat exposed in the epilogue;
//   * outputs go straight from registers as sector-complete 256-bit stores (fire and forget).
// Optional operand sharing by TMA multicast (FS2_LN_MG = 2 / 4, FS2_LN_AMC = 1): clusters of 2 MG CTAs = MG row tiles x 2
// column halves, where the two CTAs of a row tile each fetch 64 of its 128 A rows for both and the MG CTAs of a column
// half each fetch 1/MG of the weight rows for all of them; a stage is refilled only when every CTA that reads what this
// CTA writes has consumed it (the MMA warp's tcgen05.commit arrives, multicast, on the empty barriers of its row mate
// and its column mates).  Correct (tested in every configuration) but measured SLOWER than plain pairs -- the lock step
// of 4 - 8 CTAs costs more than the halved L2 -> SM traffic saves -- so the default is MG = 1 (pairs; the pair shares its A
// tile by multicast, FS2_LN_AMC = 1, which measured 5 - 10 % faster than separate fetches).
// Where the time goes (experiments with a build that could switch each part off, profiles/r02_ln_time_debug*.log; the
// switches cost registers -- spills -- in the epilogue and were removed again; 3xF16 out-projection): removing the output stores
// -31 %, the residual -15 %, the statistics exchange -13 %, all operand loads -30 % of what is left: the row-per-thread
// 256-bit stores of three output tensors are the largest single cost.
// Warp roles (352 threads): 0..7 = epilogue (warp & 3 = TMEM lane quarter, warp / 4 = column group; the groups take
// alternate 32-column chunks), 8 = TMA producer for the operand stages, 9 = MMA issuer (+ TMEM allocation), 10 = TMA
// producer for the residual ring.
// Every mbarrier wait is bounded (tc_common.cuh): a protocol bug traps instead of hanging the GPU.
#include <stdlib.h>

#include "tc_common.cuh"

namespace fs2 {
namespace {
using namespace tc;

constexpr int BM = 128;
constexpr int CHUNK_BYTES = BM * 128;                  // residual chunk: 128 rows x 32 fp32
constexpr int CL_THREADS = 352;

// KH ("half-depth K"): a stage holds 32 instead of 64 fp16 of K per row (64-byte swizzle rows).  Same bytes in flight, twice
// the stages: the 3xF16 C = 384 stage is 80 KB at K = 64, so only two fit and the ring cannot cover the L2 latency (each slot is
// re-requested only after 1152 cycles of MMA work on it have retired: per-stage time (R + M) / 2 with refill time R ~ 3000
// cycles); four 40 KB stages bring that to max(M, (R' + M) / 4).
template <int C, bool X3, bool KH = false>
struct CCfg {
  static constexpr int BKE = KH ? 32 : 64;              // fp16 K elements per stage = one swizzle row
  static constexpr int ROWB = BKE * 2;                  // bytes per operand row in a stage (128- or 64-byte swizzle)
  static constexpr int A_BYTES = BM * ROWB;             // 16 (8) KB per A plane per stage
  static constexpr int H = C / 2;                       // columns per CTA
  static constexpr int NT = H / 2;                      // values per epilogue thread
  static constexpr int NJ = NT / 32;                    // 32-column chunks per thread
  static constexpr int NCH = H / 32;                    // residual chunks per tile and CTA
  static constexpr int B_BYTES = H * ROWB;
  static constexpr int PL = X3 ? 2 : 1;
  static constexpr int STAGE_BYTES = PL * (A_BYTES + B_BYTES);
  static constexpr int A_LO = A_BYTES, B_HI = PL * A_BYTES, B_LO = B_HI + B_BYTES;
  static constexpr int VEC_BYTES = 3 * H * 4;            // bias | gamma | beta of this CTA's columns
  static constexpr int XCHG_BYTES = 2 * 4 * BM * 8;      // [slot][source][row] float2 (mean, M2)
  static constexpr int FIXED = VEC_BYTES + XCHG_BYTES + 512 /*barriers*/ + 1024 /*alignment slack*/;
  static constexpr int BUDGET = 227 * 1024 - FIXED;
  // a full tile of residual chunks when at least 2 (3xF16) / 3 (f16) operand stages still fit, else half a tile
  static constexpr int MIN_STAGES = KH ? 4 : (X3 ? 2 : 3);
  static constexpr int RB = (BUDGET - NCH * CHUNK_BYTES) / STAGE_BYTES >= MIN_STAGES ? NCH : NCH / 2;
  static constexpr int STAGES_RAW = (BUDGET - RB * CHUNK_BYTES) / STAGE_BYTES;
  static constexpr int STAGES = STAGES_RAW > 6 ? 6 : STAGES_RAW;
  static constexpr size_t SMEM = (size_t)STAGES * STAGE_BYTES + (size_t)RB * CHUNK_BYTES + FIXED;
  static constexpr uint32_t IDESC = idesc_f16(BM, H);
  static constexpr int TMEM_COLS = 2 * H <= 256 ? 256 : 512;
  static_assert(C % 128 == 0 && H % 16 == 0 && H <= 256 && NJ >= 1 && NCH % 2 == 0, "row width");
  static_assert(B_BYTES % 1024 == 0 && STAGES >= 2 && RB >= 2 && 2 * H <= 512, "resources");
};

struct ClParams {
  int M, K;
  int a_mc;                                       // A tile fetched half / half by the two CTAs of a row tile (multicast)
  int xasync;                                     // FS2_LN_XASYNC: statistics exchange by st.async + complete_tx instead of store + fence + remote arrive
  int prefetch;                                   // FS2_GEMM_PREFETCH (default 0: measured slower, see the producer loop): L2 prefetch of the next row tile's A planes and residual rows
  const float* bias; const float* gamma; const float* beta; float eps;
  int has_resid;
  float* out; int ldo;
  __half* outp; __half* outp_lo; int ldo_p;      // operand planes of the result (hi; lo for a 3xF16 consumer), nullable
  float a_inv; const float* w_inv;
};

__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release;\n\tbarrier.cluster.wait.acquire;" ::: "memory");
}
__device__ __forceinline__ uint32_t map_to_cta(uint32_t local_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void st_cluster_f32x2(uint32_t addr, float a, float b) {
  asm volatile("st.shared::cluster.v2.f32 [%0], {%1, %2};" ::"r"(addr), "f"(a), "f"(b) : "memory");
}
// asynchronous store into a peer CTA's shared memory; its 8 bytes are credited to the peer's mbarrier when they have landed
__device__ __forceinline__ void st_async_f32x2(uint32_t remote_addr, float a, float b, uint32_t remote_bar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.f32 [%0], {%1, %2}, [%3];"
               ::"r"(remote_addr), "f"(a), "f"(b), "r"(remote_bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// bounded wait with cluster-scope acquire (the peer CTA's st.shared::cluster data is visible afterwards)
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p, q;\n\t.reg .u32 cnt;\n\t"
      "mov.u32 cnt, 0;\n"
      "FS2_CWAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
      "@p bra FS2_CWAIT_DONE;\n\t"
      "add.u32 cnt, cnt, 1;\n\t"
      "setp.lt.u32 q, cnt, 4194304;\n\t"
      "@q bra FS2_CWAIT_LOOP;\n"
      "FS2_CWAIT_DONE:\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  if (!ok) {
    printf("fs2 gemm_ln cluster kernel: statistics exchange timed out (block %d thread %d)\n", blockIdx.x, threadIdx.x);
    __trap();
  }
}
// warp-convergent form (one lane elected inside the asm, tc_common.cuh)
__device__ __forceinline__ void tma_load_3d_mc_elect(uint32_t dst_addr, const CUtensorMap* map, uint32_t bar_addr, int c0, int c1, int c2, uint16_t cta_mask) {
  asm volatile(
      "{\n\t.reg .pred e;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "@e cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4, %5}], [%2], %6;\n\t}"
      ::"r"(dst_addr), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar_addr), "r"(c0), "r"(c1), "r"(c2), "h"(cta_mask)
      : "memory");
}
// MMA-completion arrive on the barrier at this offset in every CTA of `cta_mask`
__device__ __forceinline__ void tcgen05_commit_mc(uint64_t* bar, uint16_t cta_mask) {
  asm volatile(
      "{\n\t.reg .pred e;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n\t}"
      ::"r"(smem_u32(bar)), "h"(cta_mask) : "memory");
}
__device__ __forceinline__ float4 lds128(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}

template <int C, bool X3, int MG, bool KH>
__global__ void __launch_bounds__(CL_THREADS, 1)
gemm_ln_cluster_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_a64,
                       const __grid_constant__ CUtensorMap tmap_b, const __grid_constant__ CUtensorMap tmap_b_lo,
                       const __grid_constant__ CUtensorMap tmap_r, ClParams p) {
  using L = CCfg<C, X3, KH>;
  constexpr int CS = 2 * MG;                    // cluster size
  constexpr int BKE = L::BKE, ROWB = L::ROWB;
  constexpr int BROWS = L::H / MG;              // weight rows this CTA fetches per stage (for all its column mates)
  static_assert(BROWS % 8 == 0, "weight slice must keep the 8-row swizzle atom");
  extern __shared__ uint8_t smem_raw[];
  uint8_t* tiles = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* rbuf = tiles + (size_t)L::STAGES * L::STAGE_BYTES;                      // residual ring, 1024-byte aligned chunks
  float* vec = reinterpret_cast<float*>(rbuf + (size_t)L::RB * CHUNK_BYTES);       // bias | gamma | beta
  float2* xchg = reinterpret_cast<float2*>(reinterpret_cast<uint8_t*>(vec) + L::VEC_BYTES);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(xchg) + L::XCHG_BYTES);
  uint64_t* empty_bar = full_bar + 6;
  uint64_t* acc_full = empty_bar + 6;        // [2]
  uint64_t* acc_empty = acc_full + 2;        // [2]
  uint64_t* r_full = acc_empty + 2;          // [RB <= 6]
  uint64_t* r_empty = r_full + 6;            // [RB]
  uint64_t* x_bar = r_empty + 6;             // statistics exchange, 16 warp arrivals (8 local + 8 from the peer) per tile
  uint64_t* xa_bar = x_bar + 1;              // [2] st.async form of the exchange (FS2_LN_XASYNC): 8 local warps + expect_tx, by tile parity
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(xa_bar + 2);

  const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int steps = (p.K + L::BKE - 1) / L::BKE;
  const int tiles_total = (p.M + BM - 1) / BM;
  const int cluster_id = blockIdx.x / CS, n_clusters = gridDim.x / CS;
  const int nh = (int)(rank & 1u), mrow = (int)(rank >> 1);      // column half, row tile inside the cluster
  const int col0 = nh * L::H;                  // this CTA's first output column
  // CTAs that read what this CTA fetches: its row mate (A half) and the CTAs of its column half (weight slice)
  uint32_t mask_b = 0;
#pragma unroll
  for (int i = 0; i < MG; ++i) mask_b |= 1u << (2 * i + nh);
  const uint16_t mask_a = (uint16_t)(3u << (2 * mrow));
  const bool share = p.a_mc != 0 || MG > 1;
  const uint16_t mask_e = (uint16_t)((p.a_mc ? mask_a : (1u << rank)) | mask_b);
  const int n_readers = (p.a_mc ? 1 : 0) + MG;

  if (threadIdx.x == 0) {
    for (int s = 0; s < L::STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], n_readers); }
    for (int i = 0; i < 2; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], 8); }
    for (int i = 0; i < L::RB; ++i) { mbar_init(&r_full[i], 1); mbar_init(&r_empty[i], 4); }
    mbar_init(x_bar, 16);
    mbar_init(&xa_bar[0], 9); mbar_init(&xa_bar[1], 9);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 9) tmem_alloc(tmem_slot, L::TMEM_COLS);
  pdl_wait();                                  // barrier init and TMEM allocation overlapped the previous kernel's tail
  for (int i = threadIdx.x; i < L::H; i += blockDim.x) {
    vec[i] = p.bias ? __ldg(p.bias + col0 + i) : 0.f;
    vec[L::H + i] = __ldg(p.gamma + col0 + i);
    vec[2 * L::H + i] = __ldg(p.beta + col0 + i);
  }
  tcgen05_fence_before();
  __syncthreads();
  cluster_sync_all();                          // the peer's barriers exist before anything arrives on them
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 8) {
    // ---- TMA producer: operand stages.  Whole warp, one lane elected inside each asm; coordinates and addresses of a stage are
    // computed before the wait for its slot (gemm_tc.cu: what sits between "slot free" and "loads issued" is refill latency)
    const uint32_t tiles_addr = smem_u32(tiles), full_addr = smem_u32(full_bar);
    int n = 0;
    for (int rd = cluster_id; rd * MG < tiles_total; rd += n_clusters) {
      const int r0 = (rd * MG + mrow) * BM;        // a row tile past the end is all zero fill (its mates still need this CTA's share)
      // L2 prefetch of the next row tile's A planes (FS2_GEMM_PREFETCH=1; measured slower, off by default -- gemm_tc.cu)
      const int r0n = ((rd + n_clusters) * MG + mrow) * BM;
      const bool pf = p.prefetch && nh == 0 && r0n < p.M;
      for (int s = 0; s < steps; ++s, ++n) {
        if (pf && lane == 0) { tma_prefetch_3d(&tmap_a, s * BKE, r0n, 0); if (X3) tma_prefetch_3d(&tmap_a, s * BKE, r0n, 1); }
        const int slot = n % L::STAGES;
        const uint32_t st = tiles_addr + (uint32_t)slot * L::STAGE_BYTES, fb = full_addr + (uint32_t)slot * 8u;
        const int k0 = s * BKE;
        pin_before(st, fb, k0, r0);
        mbar_wait(&empty_bar[slot], ((n / L::STAGES) & 1) ^ 1);
        mbar_expect_tx_elect(fb, L::STAGE_BYTES);   // everything that lands in this stage, whoever fetches it
        if (p.a_mc) {                                // rows [64 nh, 64 nh + 64) of the A tile, for both column halves
          tma_load_3d_mc_elect(st + nh * (64 * ROWB), &tmap_a64, fb, k0, r0 + nh * 64, 0, mask_a);
          if (X3) tma_load_3d_mc_elect(st + L::A_LO + nh * (64 * ROWB), &tmap_a64, fb, k0, r0 + nh * 64, 1, mask_a);
        } else {
          tma_load_3d_elect(st, &tmap_a, fb, k0, r0, 0);
          if (X3) tma_load_3d_elect(st + L::A_LO, &tmap_a, fb, k0, r0, 1);
        }
        if (MG > 1) {                                // weight rows [BROWS mrow, +BROWS) of this column half, for every row tile
          tma_load_3d_mc_elect(st + L::B_HI + mrow * (BROWS * ROWB), &tmap_b, fb, k0, col0 + mrow * BROWS, 0, (uint16_t)mask_b);
          if (X3) tma_load_3d_mc_elect(st + L::B_LO + mrow * (BROWS * ROWB), &tmap_b_lo, fb, k0, col0 + mrow * BROWS, 0, (uint16_t)mask_b);
        } else {
          tma_load_3d_elect(st + L::B_HI, &tmap_b, fb, k0, col0, 0);
          if (X3) tma_load_3d_elect(st + L::B_LO, &tmap_b_lo, fb, k0, col0, 0);
        }
      }
    }
  } else if (warp == 9) {
    // ---- MMA issuer: whole warp, one lane elected inside each tcgen05 asm ----
    auto mkdesc = [](uint32_t a) { return KH ? make_sw64_kmajor_desc(a) : make_sw128_kmajor_desc(a); };
    int n = 0, it = 0;
    for (int rd = cluster_id; rd * MG < tiles_total; rd += n_clusters, ++it) {
      const int acc = it & 1;
      mbar_wait(&acc_empty[acc], ((it >> 1) & 1) ^ 1);
      tcgen05_fence_after();
      const uint32_t d = tmem_base + (uint32_t)(acc * L::H);
      for (int s = 0; s < steps; ++s, ++n) {
        const int slot = n % L::STAGES;
        const uint32_t base = smem_u32(tiles + (size_t)slot * L::STAGE_BYTES);   // descriptors before the wait (gemm_tc.cu)
        const uint64_t a_hi = mkdesc(base), b_hi = mkdesc(base + L::B_HI);
        const uint64_t a_lo = mkdesc(base + L::A_LO), b_lo = mkdesc(base + L::B_LO);
        pin_before64(a_hi, b_hi, a_lo, b_lo);
        mbar_wait(&full_bar[slot], (n / L::STAGES) & 1);
        tcgen05_fence_after();
#pragma unroll
        for (int k = 0; k < BKE / 16; ++k) {
          if (X3) {
            umma_f16(d, a_lo + 2 * k, b_hi + 2 * k, L::IDESC, (s | k) != 0);   // small terms first
            umma_f16(d, a_hi + 2 * k, b_lo + 2 * k, L::IDESC, 1);
            umma_f16(d, a_hi + 2 * k, b_hi + 2 * k, L::IDESC, 1);
          } else {
            umma_f16(d, a_hi + 2 * k, b_hi + 2 * k, L::IDESC, (s | k) != 0);
          }
        }
        if (share) tcgen05_commit_mc(&empty_bar[slot], mask_e);   // every CTA that fetches into this CTA's stage hears it
        else tcgen05_commit(&empty_bar[slot]);
      }
      tcgen05_commit(&acc_full[acc]);
    }
  } else if (warp == 10) {
    if (p.has_resid) {  // ---- TMA producer: residual ring (chunk c of a tile = columns col0 + 32 c .. + 31); whole warp, elected issue ----
      const uint32_t rbuf_addr = smem_u32(rbuf), rfull_addr = smem_u32(r_full);
      int q = 0;
      for (int rd = cluster_id; rd * MG < tiles_total; rd += n_clusters) {
        const int r0 = (rd * MG + mrow) * BM;
        const int r0n = ((rd + n_clusters) * MG + mrow) * BM;
        const bool pf = p.prefetch && r0n < p.M;     // FS2_GEMM_PREFETCH experiment (off)
        for (int c = 0; c < L::NCH; ++c, ++q) {
          if (pf && lane == 0) tma_prefetch_3d(&tmap_r, col0 + c * 32, r0n, 0);
          const int slot = q % L::RB;
          const uint32_t dst = rbuf_addr + (uint32_t)slot * CHUNK_BYTES, fb = rfull_addr + (uint32_t)slot * 8u;
          pin_before(dst, fb, col0 + c * 32, r0);
          mbar_wait(&r_empty[slot], ((q / L::RB) & 1) ^ 1);
          mbar_expect_tx_elect(fb, CHUNK_BYTES);
          tma_load_3d_elect(dst, &tmap_r, fb, col0 + c * 32, r0, 0);
        }
      }
    }
  } else {
    // ---- epilogue: thread == row (TMEM lane), group g takes the 32-column chunks g, g + 2, ... of this CTA's half ----
    const int wq = warp & 3, grp = warp >> 2;
    const int row = wq * 32 + lane;
    const uint32_t lane_off = (uint32_t)(wq * 32) << 16;
    const float oscale = p.a_inv * (p.w_inv ? __ldg(p.w_inv) : 1.0f);
    const uint32_t rbuf_row = smem_u32(rbuf) + (uint32_t)row * 128u;
    const uint32_t swz = (uint32_t)(row & 7);
    const uint32_t x_local = smem_u32(xchg), x_remote = map_to_cta(x_local, rank ^ 1u);
    const uint32_t xbar_remote = map_to_cta(smem_u32(x_bar), rank ^ 1u);
    const int src = nh * 2 + grp;                      // which of the four owners of a row this thread is
    const long plane = (long)p.M * p.ldo_p;
    float y[L::NT];
    int it = 0;
    for (int rd = cluster_id; rd * MG < tiles_total; rd += n_clusters, ++it) {
      const int acc = it & 1;
      const long m = (long)(rd * MG + mrow) * BM + row;
      const bool row_ok = m < p.M;
      mbar_wait(&acc_full[acc], (it >> 1) & 1);
      tcgen05_fence_after();
      const uint32_t taddr = tmem_base + lane_off + (uint32_t)(acc * L::H);
#pragma unroll
      for (int j = 0; j < L::NJ; ++j) tmem_ld32_nowait(taddr + (grp + 2 * j) * 32, y + j * 32);
      tmem_ld_wait_pin<L::NT>(y);
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[acc]);          // the accumulator is free: the next-but-one tile may start
      // y = oscale * acc + bias + resid
      float sum = 0.f;
#pragma unroll
      for (int j = 0; j < L::NJ; ++j) {
        const int c = grp + 2 * j;
        const float4* bq = reinterpret_cast<const float4*>(vec + c * 32);
        if (p.has_resid) {
          const int q = it * L::NCH + c, slot = q % L::RB;
          mbar_wait(&r_full[slot], (q / L::RB) & 1);
          const uint32_t rb = rbuf_row + (uint32_t)slot * CHUNK_BYTES;
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const float4 r4 = lds128(rb + (((uint32_t)u ^ swz) << 4));
            const float4 b4 = bq[u];
            float* yy = y + j * 32 + u * 4;
            yy[0] = fmaf(yy[0], oscale, b4.x + r4.x); yy[1] = fmaf(yy[1], oscale, b4.y + r4.y);
            yy[2] = fmaf(yy[2], oscale, b4.z + r4.z); yy[3] = fmaf(yy[3], oscale, b4.w + r4.w);
            sum += (yy[0] + yy[1]) + (yy[2] + yy[3]);
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(&r_empty[slot]);
        } else {
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const float4 b4 = bq[u];
            float* yy = y + j * 32 + u * 4;
            yy[0] = fmaf(yy[0], oscale, b4.x); yy[1] = fmaf(yy[1], oscale, b4.y);
            yy[2] = fmaf(yy[2], oscale, b4.z); yy[3] = fmaf(yy[3], oscale, b4.w);
            sum += (yy[0] + yy[1]) + (yy[2] + yy[3]);
          }
        }
      }
      // this thread's partial statistics over NT values, merged with the three other owners of the row (Chan et al.)
      const float mean_t = sum * (1.0f / L::NT);
      float m2_t = 0.f;
#pragma unroll
      for (int i = 0; i < L::NT; ++i) { const float dd = y[i] - mean_t; m2_t = fmaf(dd, dd, m2_t); }
      const uint32_t xoff = (uint32_t)((((it & 1) * 4 + src) * BM + row) * 8);
      xchg[((it & 1) * 4 + src) * BM + row] = make_float2(mean_t, m2_t);
      if (p.xasync) {
        // The peer's copy travels as an asynchronous remote store that completes on the PEER's barrier (complete_tx, like a TMA
        // load), so no cluster-scope fence is needed.  The fence of the other form (ERRBAR + CCTL.IVALL in SASS) waits for every
        // earlier memory operation of the thread -- including the previous tile's 24 output stores -- and was 10 - 14 % of this
        // kernel's stall samples.  Barriers alternate by tile parity: a CTA can be at most one tile ahead of its peer.
        const int par = it & 1;
        st_async_f32x2(x_remote + xoff, mean_t, m2_t, map_to_cta(smem_u32(&xa_bar[par]), rank ^ 1u));
        __syncwarp();
        if (lane == 0) mbar_arrive(&xa_bar[par]);                 // this warp's local partials (release at CTA scope)
        if (threadIdx.x == 0) mbar_expect_tx(&xa_bar[par], 8u * 32u * 8u);   // the peer's 256 threads x 8 bytes
        mbar_wait_cluster(&xa_bar[par], (it >> 1) & 1);
      } else {
        st_cluster_f32x2(x_remote + xoff, mean_t, m2_t);
        __syncwarp();
        if (lane == 0) {
          asm volatile("fence.acq_rel.cluster;" ::: "memory");
          mbar_arrive_cluster(map_to_cta(smem_u32(x_bar), rank));   // own barrier, same release scope
          mbar_arrive_cluster(xbar_remote);
        }
        mbar_wait_cluster(x_bar, it & 1);
      }
      float mean = 0.f, m2 = 0.f;
      float2 part[4];
#pragma unroll
      for (int s = 0; s < 4; ++s) { part[s] = xchg[((it & 1) * 4 + s) * BM + row]; mean += part[s].x; }
      mean *= 0.25f;
#pragma unroll
      for (int s = 0; s < 4; ++s) { const float dd = part[s].x - mean; m2 += part[s].y + (float)L::NT * dd * dd; }
      const float rstd = 1.0f / sqrtf(m2 * (1.0f / C) + p.eps);
      // normalise, affine, store (fp32 rows + operand planes)
#pragma unroll
      for (int j = 0; j < L::NJ; ++j) {
        const int c = grp + 2 * j;
        const float4* gq = reinterpret_cast<const float4*>(vec + L::H + c * 32);
        const float4* tq = reinterpret_cast<const float4*>(vec + 2 * L::H + c * 32);
        float* yy = y + j * 32;
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const float4 g4 = gq[u], t4 = tq[u];
          yy[u * 4 + 0] = fmaf((yy[u * 4 + 0] - mean) * rstd, g4.x, t4.x); yy[u * 4 + 1] = fmaf((yy[u * 4 + 1] - mean) * rstd, g4.y, t4.y);
          yy[u * 4 + 2] = fmaf((yy[u * 4 + 2] - mean) * rstd, g4.z, t4.z); yy[u * 4 + 3] = fmaf((yy[u * 4 + 3] - mean) * rstd, g4.w, t4.w);
        }
        if (row_ok) {
          const int col = col0 + c * 32;
          if (p.out != nullptr) {
            float* dst = p.out + m * p.ldo + col;
#pragma unroll
            for (int u = 0; u < 4; ++u) st_global_v8(dst + u * 8, yy + u * 8);
          }
          if (p.outp != nullptr) {
            uint32_t hh[16], ll[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              if (X3) split_pair(yy[2 * i], yy[2 * i + 1], hh[i], ll[i]);
              else hh[i] = hi_pair(yy[2 * i], yy[2 * i + 1]);
            }
            __half* dh = p.outp + m * p.ldo_p + col;
            st_global_v8_b32(dh, hh); st_global_v8_b32(dh + 16, hh + 8);
            if (X3 && p.outp_lo != nullptr) { st_global_v8_b32(dh + plane, ll); st_global_v8_b32(dh + plane + 16, ll + 8); }
          }
        }
      }
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  cluster_sync_all();                          // no CTA leaves while its peer may still write into it
  pdl_trigger();
  if (warp == 9) {
    tcgen05_fence_after();
    tmem_dealloc(tmem_base, L::TMEM_COLS);
  }
}

template <int C, bool X3, int MG, bool KH = false>
int launch_cl(const TapGemm& g, bool a_mc, cudaStream_t st) {
  using L = CCfg<C, X3, KH>;
  constexpr int CS = 2 * MG;
  const uint64_t M = (uint64_t)g.B * g.L;
  static unsigned long long configured = 0;   // per-device bit mask
  int rc;
  if ((rc = ensure_smem_attr(gemm_ln_cluster_kernel<C, X3, MG, KH>, L::SMEM, &configured))) return rc;
  CUtensorMap ma, ma64, mb, mb_lo, mr;
  const uint64_t arow = (uint64_t)g.K * 2;
  if ((rc = make_map(&ma, g.xp, g.K, M, X3 ? 2 : 1, arow, arow * M, BM, true, KH))) return rc;
  if ((rc = make_map(&ma64, g.xp, g.K, M, X3 ? 2 : 1, arow, arow * M, 64, true, KH))) return rc;
  if ((rc = make_map(&mb, g.w_hi, g.K, C, 1, arow, arow * C, L::H / MG, true, KH))) return rc;
  if ((rc = make_map(&mb_lo, X3 ? g.w_lo : g.w_hi, g.K, C, 1, arow, arow * C, L::H / MG, true, KH))) return rc;
  if (g.resid) { if ((rc = make_map(&mr, g.resid, C, M, 1, (uint64_t)g.ldr * 4, (uint64_t)g.ldr * 4 * M, BM, false))) return rc; }
  else mr = ma;
  ClParams p;
  p.M = (int)M; p.K = g.K; p.bias = g.bias; p.gamma = g.ln_gamma; p.beta = g.ln_beta; p.eps = g.ln_eps;
  p.a_mc = (a_mc || MG > 1) ? 1 : 0;
  p.has_resid = g.resid != nullptr;
  { static int xa = -1; if (xa < 0) { const char* e = getenv("FS2_LN_XASYNC"); xa = e ? atoi(e) : 0; } p.xasync = xa; }
  { static int pfe = -1; if (pfe < 0) { const char* e = getenv("FS2_GEMM_PREFETCH"); pfe = e ? atoi(e) : 0; } p.prefetch = pfe; }
  p.out = g.out; p.ldo = g.ldo;
  p.outp = g.outp; p.ldo_p = g.ldo_p; p.outp_lo = (g.outp && g.outp_lo) ? g.outp + (long)M * g.ldo_p : nullptr;
  p.a_inv = g.a_inv; p.w_inv = g.w_inv;
  const int tiles = (int)((M + BM - 1) / BM);
  const int rounds = (tiles + MG - 1) / MG;
  const int slots = sm_count_current() / CS;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(CS * (rounds < slots ? rounds : slots)), 1, 1);
  cfg.blockDim = dim3(CL_THREADS, 1, 1);
  cfg.dynamicSmemBytes = L::SMEM;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CS; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = pdl_enabled() ? 2 : 1;
  FS2_CUDA_CHECK(cudaLaunchKernelEx(&cfg, gemm_ln_cluster_kernel<C, X3, MG, KH>, ma, ma64, mb, mb_lo, mr, p));
  FS2_LAUNCH_CHECK();
  return FS2_OK;
}

// FS2_LN_MG: row tiles per cluster (1, 2 or 4 -> clusters of 2, 4 or 8 CTAs); FS2_LN_AMC=1: with MG = 1, the pair shares its A tile.
// Measured (c2, 3xF16, same box, gpurun_out/bench_t_*.json): MG = 1 0.34 / 0.56 ms per step for out-projection / w_2 (0.38 / 0.60
// without A sharing), MG = 2 0.50 / 0.79, MG = 4 0.50 / 0.82.  The lock step of 4 - 8 CTAs costs more than the halved operand
// traffic saves: default MG = 1 with the pair sharing its A tile; the variants stay selectable (and tested) for the record.
int ln_mg() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("FS2_LN_MG"); v = e ? atoi(e) : 1; if (v != 1 && v != 2 && v != 4) v = 1; }
  return v;
}
bool ln_amc() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("FS2_LN_AMC"); v = e ? atoi(e) : 1; }
  return v != 0;
}
// FS2_LN_KH: half-depth K stages (CCfg) for the 3xF16 C = 384 pairs, where only two full-depth stages fit: 1 = long-K launches
// (K >= 512: w_2) only (default: w_2 0.58 -> 0.53 ms per c2 step, the K = 384 out-projection is epilogue-bound and unchanged),
// 2 = every launch, 0 = off.
int ln_kh() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("FS2_LN_KH"); v = e ? atoi(e) : 1; }
  return v;
}
template <int C, bool X3>
int launch_any(const TapGemm& g, cudaStream_t st) {
  const int mg = ln_mg();
  if (mg == 4) return launch_cl<C, X3, 4>(g, true, st);
  if (mg == 2) return launch_cl<C, X3, 2>(g, true, st);
  if constexpr (X3 && C == 384) {
    const int kh = ln_kh();
    if (kh == 2 || (kh == 1 && g.K >= 512)) return launch_cl<C, X3, 1, true>(g, ln_amc(), st);
  }
  return launch_cl<C, X3, 1>(g, ln_amc(), st);
}
}  // namespace

bool gemm_ln_planes_supported(const TapGemm& g) {
  return g.taps == 1 && (g.N == 384 || g.N == 256) && g.K % 8 == 0 && g.ln_gamma && g.xp && !g.vtp && !g.vt_out && g.act == ACT_NONE;
}

int gemm_ln_planes(const TapGemm& g, cudaStream_t st) {
  FS2_REQUIRE(gemm_ln_planes_supported(g), "gemm_ln_planes: unsupported shape (N=%d taps=%d)", g.N, g.taps);
  FS2_REQUIRE(g.w_hi && (!g.precise || g.w_lo) && (g.out || g.outp), "gemm_ln_planes: operand planes / an output missing");
  FS2_REQUIRE((reinterpret_cast<uintptr_t>(g.xp) & 15) == 0 && (reinterpret_cast<uintptr_t>(g.w_hi) & 15) == 0,
              "gemm_ln_planes: operands must be 16-byte aligned");
  FS2_REQUIRE(!g.out || (g.ldo % 8 == 0 && (reinterpret_cast<uintptr_t>(g.out) & 31) == 0), "gemm_ln_planes: output rows must be 32-byte aligned");
  FS2_REQUIRE(!g.resid || (g.ldr % 4 == 0 && (reinterpret_cast<uintptr_t>(g.resid) & 15) == 0), "gemm_ln_planes: residual rows must be 16-byte aligned");
  FS2_REQUIRE(!g.outp || (g.ldo_p % 16 == 0 && (reinterpret_cast<uintptr_t>(g.outp) & 31) == 0 && (((long)g.B * g.L * g.ldo_p) % 16) == 0),
              "gemm_ln_planes: output plane rows must be 32-byte aligned");
  if ((uint64_t)g.B * g.L == 0) return FS2_OK;
  if (g.N == 384) return g.precise ? launch_any<384, true>(g, st) : launch_any<384, false>(g, st);
  return g.precise ? launch_any<256, true>(g, st) : launch_any<256, false>(g, st);
}

}  // namespace fs2
