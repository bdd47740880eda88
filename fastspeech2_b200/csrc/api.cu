// C ABI of libfs2b200.so (declared in include/fs2_b200.h): handle, checkpoint repacking and
// the host-side sequencing of the kernels for each stage of FeedForwardTransformer._forward
// (fastspeech.py:169-243).  No kernel lives here.
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <unordered_map>
#include <vector>

#include "common.cuh"

namespace fs2 {

static thread_local char g_err[512] = "";
unsigned long long g_kernel_launches = 0;
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

struct Dense {           // one Linear / Conv1d in kernel layout
  const float* w = nullptr;     // [taps][N][K]
  const __half* w_h = nullptr;  // fp16 copy (decoder side, FS2_MATH_F16)
  const __half* w_hi_h = nullptr;  // fp16 hi / lo split (3xF16, the error-compensated family)
  const __half* w_lo_h = nullptr;
  const float* bias = nullptr;  // [N] or nullptr
  int N = 0, K = 0, taps = 1;
};
struct Norm { const float* g = nullptr; const float* b = nullptr; float eps = 1e-5f; };
struct Block { Dense qkv, out, w1, w2; Norm ln1, ln2; };
struct Predictor { Dense conv[4]; Norm ln[4]; const float* head_w = nullptr; const float* head_b = nullptr; int layers = 0; };


// ---- per-kernel-class CUDA-event profiler (bench.py roofline; off by default) -------------
enum ProfClass {
  P_EMBED, P_ENC_QKV, P_ENC_ATTN, P_ENC_OUT, P_ENC_W1, P_ENC_W2, P_PRED_GEMM, P_ROWNORM, P_VAR_EMBED, P_DEC_IN, P_DEC_QKV,
  P_DEC_ATTN, P_DEC_OUT, P_DEC_W1, P_DEC_W2, P_FEAT_OUT, P_POSTNET, P_COUNT
};
static const char* kProfLabels[P_COUNT] = {
  "embed_posenc", "enc.qkv_proj", "enc.attention", "enc.out_proj", "enc.ffn_w1_conv9", "enc.ffn_w2", "predictor.tap_gemm", "row_norm",
  "variance_embed_add", "dec.embed_linear", "dec.qkv_proj", "dec.attention", "dec.out_proj", "dec.ffn_w1_conv9", "dec.ffn_w2", "feat_out", "postnet.conv5"};
struct ProfRec { int cls; cudaEvent_t a, b; double flop, bytes; };
struct Profiler {
  bool on = false;
  std::vector<ProfRec> recs;
};
static thread_local Profiler* t_prof = nullptr;
static thread_local __half* t_split_ws = nullptr;   // scratch of the running stage for the 3xF16 activation planes
static thread_local const float* t_split_of = nullptr;   // fp32 tensor whose planes t_split_ws currently holds (or null)
static thread_local int64_t t_split_rows = 0; static thread_local int t_split_K = 0;
struct ProfScope {
  ProfRec r; bool live; cudaStream_t st;
  ProfScope(int cls, double flop, double bytes, cudaStream_t s) : live(t_prof && t_prof->on), st(s) {
    if (!live) return;
    r.cls = cls; r.flop = flop; r.bytes = bytes;
    cudaEventCreate(&r.a); cudaEventCreate(&r.b);
    cudaEventRecord(r.a, st);
  }
  ~ProfScope() {
    if (!live) return;
    cudaEventRecord(r.b, st);
    t_prof->recs.push_back(r);
  }
};

}  // namespace fs2

struct fs2_handle {
  fs2::Profiler prof;
  fs2_config cfg;
  int device = 0;
  bool loaded = false;
  float* arena = nullptr;     // packed weights (owned)
  size_t arena_floats = 0;
  // encoder side
  const float* emb = nullptr; const float* enc_pe = nullptr; const float* enc_alpha = nullptr;
  std::vector<fs2::Block> enc, dec;
  fs2::Predictor dur, energy, pitch;
  const float* e_bins = nullptr; const float* p_bins = nullptr;
  const float* e_tab = nullptr; const float* e_tab_bias = nullptr;
  const float* p_tab = nullptr; const float* p_tab_bias = nullptr;
  // decoder side
  fs2::Dense dec_in; fs2::Norm dec_in_ln; const float* dec_pe = nullptr; const float* dec_alpha = nullptr;
  fs2::Dense feat_out;
  std::vector<fs2::Dense> postnet;
};

namespace fs2 {
namespace {

struct Bump {  // bump allocator over a caller-provided (or null = counting) buffer
  char* base; size_t off = 0, cap;
  Bump(void* b, size_t c) : base((char*)b), cap(c) {}
  float* floats(size_t n) { return (float*)bytes(n * sizeof(float)); }
  void* bytes(size_t n) {
    size_t a = (off + 255) & ~(size_t)255;
    off = a + n;
    return base ? base + a : nullptr;
  }
  bool ok() const { return base == nullptr || off <= cap; }
};

using Map = std::unordered_map<std::string, const fs2_weight_desc*>;

// FS2_FUSED_SPLIT=0 (debug / A-B): every 3xF16 GEMM runs its own pre-pass
bool fused_split() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("FS2_FUSED_SPLIT"); v = e ? atoi(e) : 1; }
  return v != 0;
}
int norm_rows(const RowNorm& r, cudaStream_t st);

int dense(const TapGemm& g, int math_mode, cudaStream_t st, int cls) {
  const double M = (double)g.B * g.L;
  // algorithmic bytes: operands at the width this launch reads them (fp16 copies: 2 B; fp32, or fp16 hi + lo: 4 B),
  // the result at the width(s) it is written, the residual as fp32
  const double e_in = g.x_h ? 2.0 : 4.0, e_out = (g.out ? 4.0 : 0.0) + (g.out_h ? 2.0 : 0.0);
  ProfScope prof_scope(cls, 2.0 * M * g.N * g.K * g.taps,
               e_in * (M * g.K + (double)g.taps * g.N * g.K) + M * g.N * (e_out + (g.resid ? 4.0 : 0.0)), st);
  if (math_mode == FS2_MATH_TF32 && g.ln_gamma) return gemm_ln_tf32(g, st);   // fp16 operands / fp16 copy handled inside
  if (g.x_h) return tap_gemm_f16(g, st);
  if (math_mode == MATH_3XTF32) {
    // the planes of g.x may already be in the scratch: written by the LayerNorm that produced x, or by the previous
    // GEMM's pre-pass over the same x (energy / pitch predictors share their input)
    TapGemm gs = g;
    gs.split_ws = t_split_ws;
    gs.split_ready = fused_split() && t_split_of == g.x && t_split_rows == (int64_t)g.B * g.L && t_split_K == g.K && g.ldx == g.K;
    t_split_of = g.x; t_split_rows = (int64_t)g.B * g.L; t_split_K = g.K;
    if (!fused_split() || g.ldx != g.K) t_split_of = nullptr;
    return tap_gemm_3xtf32(gs, st);
  }
  return math_mode == FS2_MATH_TF32 ? tap_gemm_tf32(g, st) : tap_gemm_fp32(g, st);
}
// LayerNorm whose output feeds a 3xF16 GEMM next: write the operand planes from the same kernel
int norm_rows_split(RowNorm r, cudaStream_t st) {
  if (fused_split() && r.out && r.ldo == r.C && t_split_ws) {
    r.split_out = t_split_ws;
    t_split_of = r.out; t_split_rows = r.rows; t_split_K = r.C;
  }
  return norm_rows(r, st);
}
int norm_rows(const RowNorm& r, cudaStream_t st) {
  ProfScope prof_scope(P_ROWNORM, 8.0 * r.rows * r.C, 4.0 * r.rows * r.C * (1 + (r.resid ? 1 : 0) + (r.out ? 1 : 0)), st);
  return row_norm(r, st);
}
int attention(int math_mode, const float* qkv, const float* vt, int lpad, const int64_t* lens, int B, int L, int C, int heads,
              float* ctx, cudaStream_t st, int cls) {
  ProfScope prof_scope(cls, 4.0 * B * (double)L * L * C, 4.0 * 4.0 * B * (double)L * C, st);
  // the encoder (MATH_3XTF32) keeps the exact-fp32 attention core: 0.2 ms at c2, and its scores feed the durations
  return math_mode == FS2_MATH_TF32 ? attention_tf32(qkv, vt, lpad, lens, B, L, C, heads, ctx, st)
                                    : attention_fp32(qkv, lens, B, L, C, heads, ctx, st);
}
inline int round4(int x) { return (x + 3) & ~3; }

TapGemm make_gemm(const Dense& d, const float* x, int ldx, int B, int L, int act, const float* resid, int ldr, float* out,
                  int ldo) {
  TapGemm g;
  g.x = x; g.ldx = ldx; g.B = B; g.L = L; g.K = d.K; g.w = d.w; g.w_h = d.w_h; g.w_hi_h = d.w_hi_h; g.w_lo_h = d.w_lo_h; g.bias = d.bias; g.N = d.N; g.taps = d.taps;
  g.act = act; g.resid = resid; g.ldr = ldr; g.out = out; g.ldo = ldo;
  return g;
}

RowNorm make_norm(const Norm& n, const float* x, int ldx, int64_t rows, int C, float* out, int ldo) {
  RowNorm r;
  memset(&r, 0, sizeof(r));
  r.x = x; r.ldx = ldx; r.gamma = n.g; r.beta = n.b; r.eps = n.eps; r.rows = rows; r.C = C; r.out = out; r.ldo = ldo;
  return r;
}

// FS2_F16_PARTS (debug / bisecting): which pieces of FS2_MATH_F16 beyond the conv-FFN are on.  2 = LayerNorm-fused
// projections (gemm_ln_tc.cu) with fp16 operands / fp16 copy, 4 = mel projection + Postnet in f16, 8 = q|k|v in f16.
int f16_parts_mask() {
  static int m = -1;
  if (m < 0) { const char* e = getenv("FS2_F16_PARTS"); m = e ? atoi(e) : 14; }
  return m;
}

// FFT blocks on xin; the result lands in *result (one of the two ping-pong buffers xin / y)
int run_blocks(const std::vector<Block>& blocks, float* xin, float* yin, float* qkv, float* vt, float* ctx, float* hid,
               const int64_t* lens, int B, int L, int C, int heads, int math_mode, bool is_dec, cudaStream_t st, float** result,
               __half* xh = nullptr) {
  const int64_t rows = (int64_t)B * L;
  const int c_qkv = is_dec ? P_DEC_QKV : P_ENC_QKV, c_att = is_dec ? P_DEC_ATTN : P_ENC_ATTN;
  const int c_out = is_dec ? P_DEC_OUT : P_ENC_OUT, c_w1 = is_dec ? P_DEC_W1 : P_ENC_W1, c_w2 = is_dec ? P_DEC_W2 : P_ENC_W2;
  float *x = xin, *y = yin;
  const int f16_parts = f16_parts_mask();
  for (const Block& k : blocks) {
    int rc;
    // q | k | v projection (attention.py:48-50), one GEMM with N = 3C
    TapGemm gq = make_gemm(k.qkv, x, C, B, L, ACT_NONE, nullptr, 0, qkv, 3 * C);
    if (math_mode == FS2_MATH_TF32) {  // V third stored transposed for the tensor-core attention (gemm_tc.cu epilogue)
      gq.vt_out = vt; gq.vt_col0 = 2 * C; gq.vt_dk = C / heads; gq.vt_heads = heads; gq.vt_lpad = round4(L);
    }
    if (xh && (f16_parts & 8)) { gq.x_h = xh; gq.ldx_h = C; }   // xh holds the fp16 copy of x here (written by the LayerNorm before)
    if ((rc = dense(gq, math_mode, st, c_qkv))) return rc;
    if ((rc = attention(math_mode, qkv, vt, round4(L), lens, B, L, C, heads, ctx, st, c_att))) return rc;
    // x = LN(x + linear_out(ctx)) (attention.py:74, encoder.py:60-62); FS2_MATH_F16: also the fp16 copy for the conv-FFN
    TapGemm go = make_gemm(k.out, ctx, C, B, L, ACT_NONE, x, C, y, C);
    go.ln_gamma = k.ln1.g; go.ln_beta = k.ln1.b; go.ln_eps = k.ln1.eps;
    const bool fuse_ln = math_mode == FS2_MATH_TF32 && gemm_ln_tf32_supported(go) && (!xh || (f16_parts & 2));
    if (fuse_ln) {   // fused: result in y, swap roles
      go.out_h = xh; go.ldo_h = C;
      if ((rc = dense(go, math_mode, st, c_out))) return rc;
      float* t = x; x = y; y = t;
    } else {
      go.ln_gamma = nullptr;
      if ((rc = dense(go, math_mode, st, c_out))) return rc;
      RowNorm r1 = make_norm(k.ln1, y, C, rows, C, x, C);
      r1.out_h = xh; r1.ldo_h = C;
      if ((rc = math_mode == MATH_3XTF32 ? norm_rows_split(r1, st) : norm_rows(r1, st))) return rc;   // feeds the conv-FFN
    }
    // conv-FFN: hid = relu(conv_k(x)); x = LN(x + conv_1(hid))  (modules.py:247-248, encoder.py:64-69)
    TapGemm g1 = make_gemm(k.w1, x, C, B, L, ACT_RELU, nullptr, 0, hid, k.w1.N);
    TapGemm g2 = make_gemm(k.w2, hid, k.w1.N, B, L, ACT_NONE, x, C, y, C);
    if (xh) {   // fp16 operands: the hidden activations only ever exist as fp16 (in the same workspace slot)
      __half* hid_h = reinterpret_cast<__half*>(hid);
      g1.x_h = xh; g1.ldx_h = C; g1.out = nullptr; g1.ldo = 0; g1.out_h = hid_h; g1.ldo_h = k.w1.N;
      g2.x_h = hid_h; g2.ldx_h = k.w1.N;
    }
    if ((rc = dense(g1, math_mode, st, c_w1))) return rc;
    g2.ln_gamma = k.ln2.g; g2.ln_beta = k.ln2.b; g2.ln_eps = k.ln2.eps;
    if (math_mode == FS2_MATH_TF32 && gemm_ln_tf32_supported(g2) && (!xh || (f16_parts & 2))) {
      g2.out_h = xh; g2.ldo_h = C;             // fp16 copy of the block output: A operand of the next q|k|v / mel projection
      if ((rc = dense(g2, math_mode, st, c_w2))) return rc;
      float* t = x; x = y; y = t;
    } else {
      g2.ln_gamma = nullptr;
      if ((rc = dense(g2, math_mode, st, c_w2))) return rc;
      RowNorm r2 = make_norm(k.ln2, y, C, rows, C, x, C);
      r2.out_h = xh; r2.ldo_h = C;
      if ((rc = math_mode == MATH_3XTF32 ? norm_rows_split(r2, st) : norm_rows(r2, st))) return rc;   // feeds the next GEMM
    }
  }
  *result = x;
  return FS2_OK;
}

// conv stack + scalar head (duration_predictor.py:64-86 / variance_predictor.py:39-60); exact fp32 or error-compensated
int run_predictor(const Predictor& p, const float* x, int C, int B, int L, float* t1, float* t2, const int64_t* lens,
                  float* head_out, int64_t* dur_out, int math_mode, cudaStream_t st) {
  const int64_t rows = (int64_t)B * L;
  const float* cur = x; int curC = C;
  for (int i = 0; i < p.layers; ++i) {
    int rc;
    if ((rc = dense(make_gemm(p.conv[i], cur, curC, B, L, ACT_RELU, nullptr, 0, t1, p.conv[i].N), math_mode, st, P_PRED_GEMM))) return rc;
    RowNorm r = make_norm(p.ln[i], t1, p.conv[i].N, rows, p.conv[i].N, t2, p.conv[i].N);
    if (i == p.layers - 1) {  // last layer: only the scalar head leaves the kernel
      r.out = nullptr; r.head_w = p.head_w; r.head_b = p.head_b; r.head_out = head_out; r.dur_out = dur_out;
      r.lens = lens; r.L = L;
    }
    if ((rc = math_mode == MATH_3XTF32 ? norm_rows_split(r, st) : norm_rows(r, st))) return rc;
    cur = t2; curC = p.conv[i].N;
  }
  return FS2_OK;
}

const fs2_weight_desc* find(const Map& m, const std::string& k) {
  auto it = m.find(k);
  return it == m.end() ? nullptr : it->second;
}

#define NEED(var, key)                                                        \
  const fs2_weight_desc* var = find(m, key);                                  \
  if (!var) { set_error("fs2_load_weights: missing key '%s'", std::string(key).c_str()); return FS2_ERR_MISSING_WEIGHT; }

struct Packer {
  fs2_handle* h; const Map& m; cudaStream_t st; Bump bump; bool counting;
  Packer(fs2_handle* h_, const Map& m_, cudaStream_t s, float* base, size_t cap)
      : h(h_), m(m_), st(s), bump(base, cap), counting(base == nullptr) {}

  int copy(const std::string& key, int64_t n, const float** out) {
    NEED(d, key);
    int64_t have = 1; for (int i = 0; i < d->ndim; ++i) have *= d->shape[i];
    if (have != n) { set_error("fs2_load_weights: '%s' has %lld elements, expected %lld", key.c_str(), (long long)have, (long long)n); return FS2_ERR_INVALID; }
    float* dst = bump.floats(n);
    if (!counting) FS2_CUDA_CHECK(cudaMemcpyAsync(dst, d->data, n * sizeof(float), cudaMemcpyDeviceToDevice, st));
    *out = dst;
    return FS2_OK;
  }
  // Linear [N,K] (taps=1) or Conv1d [N,K,taps] -> [taps][N][K]
  int dense(const std::string& wkey, const std::string& bkey, int N, int K, int taps, Dense* out, const float* scale = nullptr,
            const float* shift = nullptr) {
    NEED(w, wkey);
    int64_t have = 1; for (int i = 0; i < w->ndim; ++i) have *= w->shape[i];
    if (have != (int64_t)N * K * taps) { set_error("fs2_load_weights: '%s' has %lld elements, expected %dx%dx%d", wkey.c_str(), (long long)have, N, K, taps); return FS2_ERR_INVALID; }
    float* dst = bump.floats((size_t)N * K * taps);
    if (!counting) { int rc = pack_conv_weight((const float*)w->data, N, K, taps, scale, dst, st); if (rc) return rc; }
    out->w = dst; out->N = N; out->K = K; out->taps = taps; out->bias = shift;
    if (!bkey.empty()) { int rc = copy(bkey, N, &out->bias); if (rc) return rc; }
    return FS2_OK;
  }
  // fp16 hi / lo planes for the error-compensated kernels (gemm_tc.cu, 3xF16)
  int split(Dense* d) {
    const size_t n = (size_t)d->N * d->K * d->taps;
    __half* hh = (__half*)bump.bytes(n * sizeof(__half));
    __half* lh = (__half*)bump.bytes(n * sizeof(__half));
    if (!counting) { int rc = split_f16(d->w, hh, lh, (long)n, st); if (rc) return rc; }
    d->w_hi_h = hh; d->w_lo_h = lh;
    return FS2_OK;
  }
  // fp16 copy for the f16 family (gemm_tc.cu, HALF)
  int half(Dense* d) {
    const size_t n = (size_t)d->N * d->K * d->taps;
    __half* hw = (__half*)bump.bytes(n * sizeof(__half));
    if (!counting) { int rc = to_half(d->w, hw, (long)n, st); if (rc) return rc; }
    d->w_h = hw;
    return FS2_OK;
  }
  int norm(const std::string& prefix, int C, float eps, Norm* out) {
    int rc;
    if ((rc = copy(prefix + "weight", C, &out->g))) return rc;
    if ((rc = copy(prefix + "bias", C, &out->b))) return rc;
    out->eps = eps;
    return FS2_OK;
  }
  int blocks(const std::string& prefix, int n, int C, int H, int kffn, bool precise, bool f16_ffn, std::vector<Block>* out) {
    out->assign(n, Block());
    for (int i = 0; i < n; ++i) {
      std::string p = prefix + ".encoders_." + std::to_string(i) + ".";
      Block& b = (*out)[i];
      // fused q|k|v: three [C,C] Linear weights stacked along N
      float* wq = bump.floats((size_t)3 * C * C);
      float* bq = bump.floats((size_t)3 * C);
      const char* nm[3] = {"q", "k", "v"};
      for (int j = 0; j < 3; ++j) {
        NEED(w, p + "self_attn.linear_" + nm[j] + ".weight");
        NEED(bb, p + "self_attn.linear_" + nm[j] + ".bias");
        if (!counting) {
          FS2_CUDA_CHECK(cudaMemcpyAsync(wq + (size_t)j * C * C, w->data, (size_t)C * C * sizeof(float), cudaMemcpyDeviceToDevice, st));
          FS2_CUDA_CHECK(cudaMemcpyAsync(bq + (size_t)j * C, bb->data, (size_t)C * sizeof(float), cudaMemcpyDeviceToDevice, st));
        }
      }
      b.qkv.w = wq; b.qkv.bias = bq; b.qkv.N = 3 * C; b.qkv.K = C; b.qkv.taps = 1;
      int rc;
      if ((rc = dense(p + "self_attn.linear_out.weight", p + "self_attn.linear_out.bias", C, C, 1, &b.out))) return rc;
      if ((rc = dense(p + "feed_forward.w_1.weight", p + "feed_forward.w_1.bias", H, C, kffn, &b.w1))) return rc;
      if ((rc = dense(p + "feed_forward.w_2.weight", p + "feed_forward.w_2.bias", C, H, 1, &b.w2))) return rc;
      if (precise) for (Dense* d : {&b.qkv, &b.out, &b.w1, &b.w2}) if ((rc = split(d))) return rc;
      if (f16_ffn) for (Dense* d : {&b.qkv, &b.w1, &b.w2}) if ((rc = half(d))) return rc;
      if ((rc = norm(p + "norm1.", C, 1e-5f, &b.ln1))) return rc;   // encoder.py:37-38
      if ((rc = norm(p + "norm2.", C, 1e-5f, &b.ln2))) return rc;
    }
    return FS2_OK;
  }
  int predictor(const std::string& prefix, Predictor* out) {
    const fs2_config& c = h->cfg;
    out->layers = c.pred_layers;
    for (int i = 0; i < c.pred_layers; ++i) {
      std::string p = prefix + "conv." + std::to_string(i) + ".";
      int rc;
      if ((rc = dense(p + "0.weight", p + "0.bias", c.pred_chans, i == 0 ? c.adim : c.pred_chans, c.pred_kernel, &out->conv[i]))) return rc;
      if ((rc = split(&out->conv[i]))) return rc;
      if ((rc = norm(p + "2.layer_norm.", c.pred_chans, 1e-12f, &out->ln[i]))) return rc;  // modules.py:115
    }
    int rc;
    if ((rc = copy(prefix + "linear.weight", c.pred_chans, &out->head_w))) return rc;
    return copy(prefix + "linear.bias", 1, &out->head_b);
  }

  int run() {
    const fs2_config& c = h->cfg;
    int rc;
    // encoder (fastspeech.py:65-84)
    if ((rc = copy("encoder.embed.0.weight", (int64_t)c.idim * c.adim, &h->emb))) return rc;
    if ((rc = copy("encoder.embed.1.alpha", 1, &h->enc_alpha))) return rc;
    if ((rc = copy("encoder.embed.1.pe", (int64_t)c.pe_len * c.adim, &h->enc_pe))) return rc;
    if ((rc = blocks("encoder", c.elayers, c.adim, c.eunits, c.ffn_kernel, true, false, &h->enc))) return rc;
    if ((rc = predictor("duration_predictor.", &h->dur))) return rc;
    if ((rc = predictor("energy_predictor.predictor.", &h->energy))) return rc;
    if ((rc = predictor("pitch_predictor.predictor.", &h->pitch))) return rc;
    if ((rc = copy("energy_predictor.energy_bins", c.n_bins - 1, &h->e_bins))) return rc;
    if ((rc = copy("pitch_predictor.pitch_bins", c.n_bins - 1, &h->p_bins))) return rc;
    // energy_embed / pitch_embed: Linear(n_bins -> adim) applied to a one-hot == column gather;
    // store W^T as a [bin][channel] table (fastspeech.py:102,113,218-219)
    {
      NEED(we, "energy_embed.weight"); NEED(wp, "pitch_embed.weight");
      float* te = bump.floats((size_t)c.n_bins * c.adim);
      float* tp = bump.floats((size_t)c.n_bins * c.adim);
      if (!counting) {
        if ((rc = pack_transpose((const float*)we->data, c.adim, c.n_bins, te, st))) return rc;
        if ((rc = pack_transpose((const float*)wp->data, c.adim, c.n_bins, tp, st))) return rc;
      }
      h->e_tab = te; h->p_tab = tp;
      if ((rc = copy("energy_embed.bias", c.adim, &h->e_tab_bias))) return rc;
      if ((rc = copy("pitch_embed.bias", c.adim, &h->p_tab_bias))) return rc;
    }
    // decoder (fastspeech.py:119-136; input layer core/encoder.py:118-125)
    if ((rc = dense("decoder.embed.0.weight", "decoder.embed.0.bias", c.ddim, c.adim, 1, &h->dec_in))) return rc;
    if ((rc = split(&h->dec_in))) return rc;                                   // hi/lo copies serve FS2_MATH_3XTF32
    if ((rc = norm("decoder.embed.1.", c.ddim, 1e-5f, &h->dec_in_ln))) return rc;
    if ((rc = copy("decoder.embed.4.alpha", 1, &h->dec_alpha))) return rc;
    if ((rc = copy("decoder.embed.4.pe", (int64_t)c.pe_len * c.ddim, &h->dec_pe))) return rc;
    if ((rc = blocks("decoder", c.dlayers, c.ddim, c.dunits, c.ffn_kernel, true, true, &h->dec))) return rc;
    if ((rc = dense("feat_out.weight", "feat_out.bias", c.odim, c.ddim, 1, &h->feat_out))) return rc;
    if ((rc = split(&h->feat_out))) return rc;
    if ((rc = half(&h->feat_out))) return rc;
    // Postnet: Conv1d(no bias) + BatchNorm1d(eval) folded into weight scale + bias (modules.py:283-348)
    h->postnet.assign(c.postnet_layers, Dense());
    for (int i = 0; i < c.postnet_layers; ++i) {
      std::string p = "postnet.postnet." + std::to_string(i) + ".";
      int cin = i == 0 ? c.odim : c.postnet_chans;
      int cout = i == c.postnet_layers - 1 ? c.odim : c.postnet_chans;
      float* scale = bump.floats(cout);
      float* shift = bump.floats(cout);
      NEED(g, p + "1.weight"); NEED(b, p + "1.bias"); NEED(mu, p + "1.running_mean"); NEED(var, p + "1.running_var");
      if (!counting && (rc = fold_batchnorm((const float*)g->data, (const float*)b->data, (const float*)mu->data,
                                            (const float*)var->data, 1e-5f, cout, scale, shift, st))) return rc;
      if ((rc = dense(p + "0.weight", "", cout, cin, c.postnet_filts, &h->postnet[i], scale, shift))) return rc;
      if ((rc = split(&h->postnet[i]))) return rc;
      if ((rc = half(&h->postnet[i]))) return rc;
    }
    return FS2_OK;
  }
};

inline int64_t max_width(const fs2_config& c) {
  int64_t w = c.adim;
  for (int v : {c.ddim, c.eunits, c.dunits, c.pred_chans, c.postnet_chans, c.odim}) if (v > w) w = v;
  return w;
}
struct EncodePlan { float *x, *y, *qkv, *ctx, *hid, *t1, *t2; __half* split; };
EncodePlan plan_encode(const fs2_config& c, Bump& b, int64_t rows) {
  EncodePlan p;
  p.x = b.floats(rows * c.adim); p.y = b.floats(rows * c.adim); p.qkv = b.floats(rows * 3 * c.adim);
  p.ctx = b.floats(rows * c.adim); p.hid = b.floats(rows * c.eunits);
  p.t1 = b.floats(rows * c.pred_chans); p.t2 = b.floats(rows * c.pred_chans);
  p.split = (__half*)b.floats(rows * max_width(c));     // 3xF16 activation planes (hi + lo = the bytes of the widest fp32 operand)
  return p;
}
struct DecodePlan { float *hm2, *x, *y, *qkv, *vt, *ctx, *hid, *t1, *t2, *q1, *q2; __half *xh, *before_h, *split; };
DecodePlan plan_decode(const fs2_config& c, Bump& b, int64_t rows, int B, int L) {
  DecodePlan p;
  p.hm2 = b.floats(rows * c.adim);
  p.x = b.floats(rows * c.ddim); p.y = b.floats(rows * c.ddim); p.qkv = b.floats(rows * 3 * c.ddim);
  p.vt = b.floats((int64_t)B * c.ddim * round4(L));
  p.ctx = b.floats(rows * c.ddim); p.hid = b.floats(rows * c.dunits);
  p.t1 = b.floats(rows * c.pred_chans); p.t2 = b.floats(rows * c.pred_chans);
  p.q1 = b.floats(rows * c.postnet_chans); p.q2 = b.floats(rows * c.postnet_chans);
  p.xh = (__half*)b.bytes((size_t)rows * c.ddim * sizeof(__half));   // FS2_MATH_F16: fp16 copy of the current block input / conv-FFN input
  p.before_h = (__half*)b.bytes((size_t)rows * c.odim * sizeof(__half));   // FS2_MATH_F16: fp16 copy of before_outs for the Postnet
  p.split = (__half*)b.floats(rows * max_width(c));
  return p;
}

}  // namespace
}  // namespace fs2

using namespace fs2;

extern "C" {

const char* fs2_last_error(void) { return g_err; }
const char* fs2_version(void) { return "fs2-b200 0.1 sm_100a"; }
unsigned long long fs2_kernel_launches(void) { return g_kernel_launches; }

int fs2_create(fs2_handle** out, const fs2_config* cfg, int device) {
  FS2_REQUIRE(out && cfg, "fs2_create: null argument");
  FS2_REQUIRE(cfg->aheads > 0 && cfg->adim % cfg->aheads == 0 && cfg->ddim % cfg->aheads == 0, "fs2_create: dims not divisible by heads");
  FS2_REQUIRE((cfg->adim == 256 || cfg->adim == 384) && (cfg->ddim == 256 || cfg->ddim == 384),
              "fs2_create: adim/ddim must be 256 or 384 (got %d/%d); kernels are specialised for configs/default.yaml", cfg->adim, cfg->ddim);
  FS2_REQUIRE(cfg->pred_chans == 256 || cfg->pred_chans == 384, "fs2_create: predictor channels must be 256 or 384");
  FS2_REQUIRE(cfg->pred_layers >= 1 && cfg->pred_layers <= 4, "fs2_create: 1..4 predictor layers");
  FS2_REQUIRE(cfg->eunits % 16 == 0 && cfg->dunits % 16 == 0 && cfg->odim % 16 == 0 && cfg->postnet_chans % 16 == 0, "fs2_create: channel counts must be multiples of 16");
  FS2_REQUIRE((cfg->ffn_kernel & 1) && (cfg->pred_kernel & 1) && (cfg->postnet_filts & 1), "fs2_create: kernel sizes must be odd");
  FS2_REQUIRE(cfg->postnet_layers >= 1, "fs2_create: postnet_layers == 0 is not supported");
  FS2_REQUIRE(cfg->n_bins % 4 == 0, "fs2_create: n_bins must be a multiple of 4");
  FS2_REQUIRE(cfg->math_mode >= FS2_MATH_FP32 && cfg->math_mode <= FS2_MATH_F16, "fs2_create: bad math_mode");
  FS2_CUDA_CHECK(cudaSetDevice(device));
  fs2_handle* h = new fs2_handle();
  h->cfg = *cfg;
  h->device = device;
  *out = h;
  return FS2_OK;
}

void fs2_destroy(fs2_handle* h) {
  if (!h) return;
  if (h->arena) cudaFree(h->arena);
  delete h;
}

int fs2_profile_enable(fs2_handle* h, int on) {
  FS2_REQUIRE(h, "fs2_profile_enable: null handle");
  h->prof.on = on != 0;
  return FS2_OK;
}
int fs2_profile_classes(void) { return P_COUNT; }
const char* fs2_profile_label(int i) { return i >= 0 && i < P_COUNT ? kProfLabels[i] : ""; }
int fs2_profile_read(fs2_handle* h, double* ms, int64_t* launches, double* flop, double* bytes) {
  FS2_REQUIRE(h && ms && launches && flop && bytes, "fs2_profile_read: null argument");
  for (int i = 0; i < P_COUNT; ++i) { ms[i] = 0; launches[i] = 0; flop[i] = 0; bytes[i] = 0; }
  for (ProfRec& r : h->prof.recs) {
    FS2_CUDA_CHECK(cudaEventSynchronize(r.b));
    float t = 0.f;
    FS2_CUDA_CHECK(cudaEventElapsedTime(&t, r.a, r.b));
    ms[r.cls] += t; launches[r.cls] += 1; flop[r.cls] += r.flop; bytes[r.cls] += r.bytes;
    cudaEventDestroy(r.a); cudaEventDestroy(r.b);
  }
  h->prof.recs.clear();
  return FS2_OK;
}

int fs2_set_math_mode(fs2_handle* h, int math_mode) {
  FS2_REQUIRE(h, "fs2_set_math_mode: null handle");
  FS2_REQUIRE(math_mode >= FS2_MATH_FP32 && math_mode <= FS2_MATH_F16, "fs2_set_math_mode: bad mode %d", math_mode);
  h->cfg.math_mode = math_mode;
  return FS2_OK;
}

int fs2_load_weights(fs2_handle* h, const fs2_weight_desc* w, int n, void* stream) {
  FS2_REQUIRE(h && w && n > 0, "fs2_load_weights: null argument");
  cudaStream_t st = (cudaStream_t)stream;
  FS2_CUDA_CHECK(cudaSetDevice(h->device));
  Map m;
  for (int i = 0; i < n; ++i) {
    FS2_REQUIRE(w[i].name && w[i].data, "fs2_load_weights: entry %d has a null name/data", i);
    m[w[i].name] = &w[i];
  }
  h->loaded = false;
  size_t need;
  {
    Packer count(h, m, st, nullptr, 0);
    int rc = count.run();
    if (rc) return rc;
    need = count.bump.off + 256;
  }
  if (need > h->arena_floats * sizeof(float)) {
    if (h->arena) { FS2_CUDA_CHECK(cudaStreamSynchronize(st)); FS2_CUDA_CHECK(cudaFree(h->arena)); h->arena = nullptr; }
    FS2_CUDA_CHECK(cudaMalloc(&h->arena, need));
    h->arena_floats = need / sizeof(float);
  }
  Packer pack(h, m, st, h->arena, need);
  int rc = pack.run();
  if (rc) return rc;
  h->loaded = true;
  return FS2_OK;
}

int fs2_workspace_bytes(fs2_handle* h, int B, int Tmax, int Lmax, size_t* out) {
  FS2_REQUIRE(h && out && B >= 0 && Tmax >= 0 && Lmax >= 0, "fs2_workspace_bytes: bad argument");
  Bump e(nullptr, 0), d(nullptr, 0);
  plan_encode(h->cfg, e, (int64_t)B * Tmax);
  plan_decode(h->cfg, d, (int64_t)B * Lmax, B, Lmax);
  *out = (e.off > d.off ? e.off : d.off) + 1024;
  return FS2_OK;
}

int fs2_encode(fs2_handle* h, const int64_t* xs, const int64_t* ilens, int B, int Tmax, float* hs, float* d_log,
               int64_t* d_int, void* ws, size_t ws_bytes, void* stream) {
  FS2_REQUIRE(h && xs && ilens && hs && ws, "fs2_encode: null argument");
  if (!h->loaded) { set_error("fs2_encode: weights not loaded"); return FS2_ERR_NOT_LOADED; }
  FS2_REQUIRE(Tmax <= h->cfg.pe_len, "fs2_encode: Tmax=%d exceeds the positional table (%d rows)", Tmax, h->cfg.pe_len);
  cudaStream_t st = (cudaStream_t)stream;
  const fs2_config& c = h->cfg;
  t_prof = &h->prof;
  Bump b(ws, ws_bytes);
  EncodePlan p = plan_encode(c, b, (int64_t)B * Tmax);
  if (!b.ok()) { set_error("fs2_encode: workspace too small (%zu < %zu)", ws_bytes, b.off); return FS2_ERR_WORKSPACE; }
  t_split_ws = p.split; t_split_of = nullptr;
  int rc;
  // the encoder's output feeds round() in the duration predictor: exact fp32 FMA in FS2_MATH_FP32,
  // error-compensated 3xF16 on the tensor cores in every other mode (never a plain 10-bit-mantissa product)
  const int precise = c.math_mode == FS2_MATH_FP32 ? FS2_MATH_FP32 : MATH_3XTF32;
  { ProfScope prof_scope(P_EMBED, 0, 8.0 * B * Tmax * c.adim, st);
    if ((rc = embed_posenc(xs, h->emb, c.idim, h->enc_pe, h->enc_alpha, B, Tmax, c.adim, p.x, st))) return rc; }
  float* enc_out = nullptr;
  if ((rc = run_blocks(h->enc, p.x, p.y, p.qkv, nullptr, p.ctx, p.hid, ilens, B, Tmax, c.adim, c.aheads, precise, false, st, &enc_out))) return rc;
  FS2_CUDA_CHECK(cudaMemcpyAsync(hs, enc_out, (size_t)B * Tmax * c.adim * sizeof(float), cudaMemcpyDeviceToDevice, st));
  if (d_log || d_int)
    if ((rc = run_predictor(h->dur, enc_out, c.adim, B, Tmax, p.t1, p.t2, ilens, d_log, d_int, precise, st))) return rc;
  return FS2_OK;
}

int fs2_length_plan(void* ds, int ds_dtype, const int64_t* ilens, float alpha, int B, int Tmax, int mutate_ds,
                    int32_t* cum, int64_t* olens, int64_t* stats, void* stream) {
  FS2_REQUIRE(ds && ilens && cum && olens && stats, "fs2_length_plan: null argument");
  return length_plan(ds, ds_dtype, ilens, alpha, B, Tmax, mutate_ds, cum, olens, stats, (cudaStream_t)stream);
}

int fs2_length_gather(const float* hs, const int32_t* cum, const int64_t* ilens, int B, int Tmax, int C, float* out,
                      int Lcap, void* stream) {
  FS2_REQUIRE(hs && cum && ilens && (out || Lcap == 0), "fs2_length_gather: null argument");
  return length_gather(hs, cum, ilens, B, Tmax, C, out, Lcap, (cudaStream_t)stream);
}

int fs2_decode(fs2_handle* h, const float* hm, const int64_t* olens, const float* es, const float* ps, int B, int L,
               float* before, float* after, float* e_out, float* p_out, int64_t* e_ids, int64_t* p_ids, void* ws,
               size_t ws_bytes, void* stream) {
  FS2_REQUIRE(h && hm && before && after && e_out && p_out && ws, "fs2_decode: null argument");
  FS2_REQUIRE((es == nullptr) == (ps == nullptr), "fs2_decode: es and ps must both be given or both be NULL");
  if (!h->loaded) { set_error("fs2_decode: weights not loaded"); return FS2_ERR_NOT_LOADED; }
  FS2_REQUIRE(L <= h->cfg.pe_len, "fs2_decode: L=%d exceeds the positional table (%d rows)", L, h->cfg.pe_len);
  cudaStream_t st = (cudaStream_t)stream;
  const fs2_config& c = h->cfg;
  const int64_t rows = (int64_t)B * L;
  t_prof = &h->prof;
  Bump b(ws, ws_bytes);
  DecodePlan p = plan_decode(c, b, rows, B, L);
  if (!b.ok()) { set_error("fs2_decode: workspace too small (%zu < %zu)", ws_bytes, b.off); return FS2_ERR_WORKSPACE; }
  t_split_ws = p.split; t_split_of = nullptr;
  const bool f16_ffn = c.math_mode == FS2_MATH_F16;                     // tf32 everywhere except the conv-FFN
  const int mode = f16_ffn ? FS2_MATH_TF32 : c.math_mode;
  const int precise = mode == FS2_MATH_FP32 ? FS2_MATH_FP32 : MATH_3XTF32;
  int rc;
  // energy / pitch predictors on the length-regulated states (fastspeech.py:195-196,214-216); fp32-class
  if ((rc = run_predictor(h->energy, hm, c.adim, B, L, p.t1, p.t2, olens, e_out, nullptr, precise, st))) return rc;
  if ((rc = run_predictor(h->pitch, hm, c.adim, B, L, p.t1, p.t2, olens, p_out, nullptr, precise, st))) return rc;
  // hs + pitch_embed(one_hot) + energy_embed(one_hot) (fastspeech.py:218-219)
  { ProfScope prof_scope(P_VAR_EMBED, 0, 4.0 * rows * c.adim * 4, st);
  if ((rc = variance_embed_add(hm, es ? es : e_out, ps ? ps : p_out, h->e_bins, h->p_bins, c.n_bins - 1, h->e_tab,
                               h->e_tab_bias, h->p_tab, h->p_tab_bias, rows, c.adim, p.hm2, e_ids, p_ids, st))) return rc; }
  // decoder input layer: Linear -> LayerNorm -> ReLU -> x + alpha*pe (core/encoder.py:118-125)
  if ((rc = dense(make_gemm(h->dec_in, p.hm2, c.adim, B, L, ACT_NONE, nullptr, 0, p.y, c.ddim), mode, st, P_DEC_IN))) return rc;
  {
    RowNorm r = make_norm(h->dec_in_ln, p.y, c.ddim, rows, c.ddim, p.x, c.ddim);
    r.relu_after = 1; r.pe = h->dec_pe; r.alpha = h->dec_alpha; r.L = L;
    if (f16_ffn) { r.out_h = p.xh; r.ldo_h = c.ddim; }    // first block's q|k|v reads the fp16 copy
    if ((rc = mode == MATH_3XTF32 ? norm_rows_split(r, st) : norm_rows(r, st))) return rc;
  }
  float* dec_out = nullptr;
  if ((rc = run_blocks(h->dec, p.x, p.y, p.qkv, p.vt, p.ctx, p.hid, olens, B, L, c.ddim, c.aheads, mode, true, st, &dec_out, f16_ffn ? p.xh : nullptr))) return rc;
  // mel linear (fastspeech.py:228-230); FS2_MATH_F16: from the fp16 copy of the last block's output, and the Postnet
  // chain stays in fp16 until the final residual layer
  const bool f16_post = f16_ffn && (f16_parts_mask() & 4) && c.odim % 16 == 0;
  {
    TapGemm g = make_gemm(h->feat_out, dec_out, c.ddim, B, L, ACT_NONE, nullptr, 0, before, c.odim);
    if (f16_post) { g.x_h = p.xh; g.ldx_h = c.ddim; g.out_h = p.before_h; g.ldo_h = c.odim; }
    if ((rc = dense(g, mode, st, P_FEAT_OUT))) return rc;
  }
  // Postnet + residual (fastspeech.py:236-238, modules.py:350-359)
  const float* cur = before; int curC = c.odim;
  const __half* cur_h = p.before_h;
  float* pp[2] = {p.q1, p.q2};
  for (int i = 0; i < c.postnet_layers; ++i) {
    bool last = i == c.postnet_layers - 1;
    float* dst = last ? after : pp[i & 1];
    TapGemm g = make_gemm(h->postnet[i], cur, curC, B, L, last ? ACT_NONE : ACT_TANH, last ? before : nullptr, c.odim, dst,
                          h->postnet[i].N);
    if (f16_post) {
      g.x_h = cur_h; g.ldx_h = curC;
      if (!last) { g.out = nullptr; g.ldo = 0; g.out_h = reinterpret_cast<__half*>(dst); g.ldo_h = h->postnet[i].N; }
      cur_h = reinterpret_cast<const __half*>(dst);
    }
    if ((rc = dense(g, mode, st, P_POSTNET))) return rc;
    cur = dst; curC = h->postnet[i].N;
  }
  return FS2_OK;
}

int fs2_masked_losses(const float* before, const float* after, const float* ys, int ld_ys_time, const float* d_out,
                      const void* ds, int ds_dtype, const float* e_out, const float* p_out, const float* es,
                      const float* ps, const int64_t* ilens, const int64_t* olens, int B, int Tmax, int L, int odim,
                      float* out7, void* scratch, void* stream) {
  FS2_REQUIRE(before && after && ys && d_out && ds && e_out && p_out && es && ps && ilens && olens && out7 && scratch,
              "fs2_masked_losses: null argument");
  return masked_losses(before, after, ys, ld_ys_time, d_out, ds, ds_dtype, e_out, p_out, es, ps, ilens, olens, B, Tmax, L,
                       odim, out7, scratch, (cudaStream_t)stream);
}

int fs2_bucketize(const float* vals, const float* bins, int n_edges, int64_t n, int64_t* ids, void* stream) {
  FS2_REQUIRE(vals && bins && ids, "fs2_bucketize: null argument");
  return bucketize(vals, bins, n_edges, n, ids, (cudaStream_t)stream);
}
int fs2_one_hot(const int64_t* ids, int64_t n, int n_bins, float* out, void* stream) {
  FS2_REQUIRE(ids && out, "fs2_one_hot: null argument");
  return one_hot(ids, n, n_bins, out, (cudaStream_t)stream);
}
int fs2_op_tap_gemm(int math_mode, const float* x, int B, int L, int K, const float* w, const float* bias, int N, int taps,
                    int act, const float* resid, float* out, void* stream) {
  FS2_REQUIRE(x && w && out, "fs2_op_tap_gemm: null argument");
  Dense d; d.w = w; d.bias = bias; d.N = N; d.K = K; d.taps = taps;
  cudaStream_t st = (cudaStream_t)stream;
  if (math_mode == FS2_MATH_F16) {   // single-operator entry for the f16 family (tests): fp16 copies made on the fly
    const size_t nx = (size_t)B * L * K, nw = (size_t)N * K * taps;
    __half* tmp = nullptr;
    FS2_CUDA_CHECK(cudaMallocAsync(&tmp, (nx + nw + 16) * sizeof(__half), st));
    __half* wh = tmp + ((nx + 7) & ~(size_t)7);
    int rc = to_half(x, tmp, (long)nx, st);
    if (!rc) rc = to_half(w, wh, (long)nw, st);
    TapGemm g = make_gemm(d, x, K, B, L, act, resid, N, out, N);
    g.x_h = tmp; g.ldx_h = K; g.w_h = wh;
    if (!rc) rc = dense(g, FS2_MATH_TF32, st, P_DEC_W1);
    cudaFreeAsync(tmp, st);
    return rc;
  }
  if (math_mode != MATH_3XTF32) return dense(make_gemm(d, x, K, B, L, act, resid, N, out, N), math_mode, st, P_DEC_W1);
  // single-operator entry for the error-compensated family (tests): split the weights on the fly
  const size_t n = (size_t)N * K * taps, nx = (size_t)B * L * K;
  __half* tmp = nullptr;
  const size_t n8 = (n + 7) & ~(size_t)7;
  FS2_CUDA_CHECK(cudaMallocAsync(&tmp, (2 * n8 + 2 * nx + 16) * sizeof(__half), st));
  int rc = split_f16(w, tmp, tmp + n8, (long)n, st);
  d.w_hi_h = tmp; d.w_lo_h = tmp + n8;
  t_split_ws = tmp + 2 * n8; t_split_of = nullptr;
  if (!rc) rc = dense(make_gemm(d, x, K, B, L, act, resid, N, out, N), math_mode, st, P_DEC_W1);
  cudaFreeAsync(tmp, st);
  return rc;
}
int fs2_op_gemm_layernorm(const float* x, int64_t rows, int K, const float* w, const float* bias, const float* resid,
                          const float* gamma, const float* beta, float eps, float* out, void* stream) {
  FS2_REQUIRE(x && w && gamma && beta && out, "fs2_op_gemm_layernorm: null argument");
  FS2_REQUIRE(rows < (1LL << 31), "fs2_op_gemm_layernorm: too many rows");
  Dense d; d.w = w; d.bias = bias; d.N = 384; d.K = K; d.taps = 1;
  TapGemm g = make_gemm(d, x, K, 1, (int)rows, ACT_NONE, resid, 384, out, 384);
  g.ln_gamma = gamma; g.ln_beta = beta; g.ln_eps = eps;
  return dense(g, FS2_MATH_TF32, (cudaStream_t)stream, P_DEC_OUT);
}
int fs2_op_attention(int math_mode, const float* qkv, const int64_t* lens, int B, int L, int C, int heads, float* ctx,
                     void* stream) {
  FS2_REQUIRE(qkv && ctx, "fs2_op_attention: null argument");
  cudaStream_t st = (cudaStream_t)stream;
  if (math_mode != FS2_MATH_TF32) return attention(math_mode, qkv, nullptr, 0, lens, B, L, C, heads, ctx, st, P_DEC_ATTN);
  // single-operator entry (tests): build the transposed V the projection epilogue normally provides
  float* vt = nullptr;
  const int lpad = round4(L);
  FS2_CUDA_CHECK(cudaMallocAsync(&vt, (size_t)B * C * lpad * sizeof(float), st));
  int rc = transpose_v(qkv, B, L, C, heads, vt, lpad, st);
  if (!rc) rc = attention(math_mode, qkv, vt, lpad, lens, B, L, C, heads, ctx, st, P_DEC_ATTN);
  cudaFreeAsync(vt, st);
  return rc;
}
int fs2_op_layernorm(const float* x, const float* resid, const float* g, const float* b, float eps, int64_t rows, int C,
                     float* out, void* stream) {
  FS2_REQUIRE(x && g && b && out, "fs2_op_layernorm: null argument");
  Norm n; n.g = g; n.b = b; n.eps = eps;
  RowNorm r = make_norm(n, x, C, rows, C, out, C);
  r.resid = resid; r.ldr = C;
  return row_norm(r, (cudaStream_t)stream);
}

}  // extern "C"
