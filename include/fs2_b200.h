/*
 * fs2_b200.h -- C ABI of libfs2b200.so, the sm_100a FastSpeech2 mel-synthesis forward path.
 *
 * The reference (rishikksh20/FastSpeech2) has no FFI layer: its operator API for this
 * path is the Python class `FeedForwardTransformer` in fastspeech.py, whose stages call
 * torch.nn modules.  Each entry point below replaces one stage of
 * `FeedForwardTransformer._forward` (fastspeech.py:169-243) / `forward` (:245-337); the
 * reference lines a function replaces are cited at its declaration.  The Python class in
 * fastspeech2_b200/fastspeech.py binds these with ctypes (INTEGRATION.md shows the stub).
 *
 * Conventions
 *   - extern "C", plain pointers and sizes, no torch / C++ types in signatures.
 *   - every function returns 0 on success or a negative FS2_ERR_* code;
 *     fs2_last_error() returns a thread-local message for the last failure.
 *   - all tensor pointers are DEVICE pointers (fp32 row-major [batch, time, channel],
 *     lengths / ids / durations int64) unless a parameter says "host".
 *   - the caller owns every input, output and workspace buffer; the library owns only the
 *     opaque handle (packed weights, TMA descriptors, launch plans).
 *   - all work is enqueued on the `stream` argument (a cudaStream_t passed as void*);
 *     no function synchronises the device.  fs2_length_plan writes its two result words to
 *     device memory; the caller reads them back (that is the path's single host sync).
 *   - a handle is bound to one device and is not thread-safe; entry points that take a handle select its device for
 *     the call and restore the caller's.  Handle-less entry points (LengthRegulator, losses, single operators, train /
 *     STFT / peer kernels) run on the CURRENT device like any CUDA library call: select the device your buffers live on.
 */
#ifndef FS2_B200_H_
#define FS2_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FS2_OK 0
#define FS2_ERR_INVALID (-1)        /* bad argument / unsupported shape            */
#define FS2_ERR_CUDA (-2)           /* a CUDA runtime / driver call failed         */
#define FS2_ERR_MISSING_WEIGHT (-3) /* fs2_load_weights: a required key is absent  */
#define FS2_ERR_WORKSPACE (-4)      /* workspace too small                         */
#define FS2_ERR_NOT_LOADED (-5)     /* stage called before fs2_load_weights        */

/* duration dtypes accepted by the length regulator (tests/test_fastspeech2.py:17 feeds floats) */
#define FS2_DUR_I64 0
#define FS2_DUR_F32 1
#define FS2_DUR_I32 2

/* arithmetic of the dense contractions (accumulation is fp32 in TMEM / registers everywhere).
 *   FS2_MATH_FP32: fp32 FMA on CUDA cores everywhere.
 *   FS2_MATH_3XTF32 ("3xf16", the reference-precision tensor-core mode and the Python class's default): every dense
 *     contraction -- projections, convolutions AND both attention products -- error-compensated on tcgen05: each operand
 *     is held as two fp16 planes hi = rn(s x), lo = rn(s x - hi) (s an exact power of two, undone in the epilogue), and
 *     a product is accumulated as hi.hi + hi.lo + lo.hi.  fp32-class results (max-abs ~1e-5 on the mels).
 *   FS2_MATH_F16: encoder + predictors as in FS2_MATH_3XTF32 (their outputs feed round() / bucketize()); the decoder
 *     side (input Linear, q|k|v, attention, out-projection, conv-FFN, mel Linear, Postnet) on kind::f16 over the hi
 *     planes only: 10-bit-mantissa operands (like tf32, round-to-nearest), twice the tf32 MMA rate.
 *   FS2_MATH_TF32: encoder + predictors as above; decoder side on kind::tf32 reading the fp32 rows directly.
 * Normalisation, softmax statistics, gathers and every integer kernel are fp32 / exact in all modes. */
#define FS2_MATH_FP32 0
#define FS2_MATH_TF32 1
#define FS2_MATH_3XTF32 2
#define FS2_MATH_F16 3

typedef struct fs2_handle fs2_handle;

/* Shapes of the network: the subset of `hp` read by FeedForwardTransformer.__init__
 * (fastspeech.py:37-160; values in configs/default.yaml:38-106). */
typedef struct fs2_config {
  int32_t idim, odim;               /* 68 symbols, 80 mel bins                            */
  int32_t adim, aheads, elayers, eunits;   /* encoder: 256, 2, 4, 1024                    */
  int32_t ddim, dlayers, dunits;           /* decoder: 384, 4, 1024 (heads shared)        */
  int32_t ffn_kernel;               /* positionwise_conv_kernel_size: 9                   */
  int32_t pred_layers, pred_chans, pred_kernel; /* 2, 256, 3                              */
  int32_t postnet_layers, postnet_chans, postnet_filts; /* 5, 256, 5                      */
  int32_t n_bins;                   /* 256 pitch / energy buckets                         */
  int32_t pe_len;                   /* informational: the positional tables' row counts are read off the tensors at load time */
  int32_t math_mode;                /* FS2_MATH_*                                         */
} fs2_config;

/* One checkpoint tensor, addressed by its reference state_dict key. */
typedef struct fs2_weight_desc {
  const char* name;     /* e.g. "decoder.encoders_.0.feed_forward.w_1.weight"  */
  const void* data;     /* device pointer, contiguous                           */
  int32_t ndim;
  int64_t shape[4];
  int32_t dtype;        /* 0 = float32, 1 = int64                               */
} fs2_weight_desc;

const char* fs2_last_error(void);
/* "fs2-b200 <ver> sm_100a" ; lets the binding check it loaded the right library */
const char* fs2_version(void);
/* number of kernels this library has enqueued in this process (bench.py reports the per-step delta) */
unsigned long long fs2_kernel_launches(void);

/* ---- handle --------------------------------------------------------------------------- */
/* replaces FeedForwardTransformer.__init__ shape plumbing (fastspeech.py:37-160) */
int fs2_create(fs2_handle** out, const fs2_config* cfg, int device);
void fs2_destroy(fs2_handle* h);
int fs2_set_math_mode(fs2_handle* h, int math_mode);

/* Repack the checkpoint into kernel layout: Conv1d [N,K,taps] -> [taps][N][K]; q/k/v Linear
 * concatenated; BatchNorm1d (eval) folded into the Postnet convolutions
 * (core/modules.py:283-348); pitch/energy embedding Linear transposed to a [bin][channel]
 * table.  Replaces load_state_dict -> module attribute reads of the reference. */
int fs2_load_weights(fs2_handle* h, const fs2_weight_desc* w, int n, void* stream);

/* Bytes of scratch the two big stages need (max of both) for a batch of B utterances,
 * Tmax phonemes and Lmax frames. */
int fs2_workspace_bytes(fs2_handle* h, int B, int Tmax, int Lmax, size_t* out);

/* ---- per-kernel-class CUDA-event profiler (used by bench.py for the roofline; off by default) -
 * While enabled, every kernel the stage functions enqueue is bracketed by two events on the
 * launch stream.  fs2_profile_read synchronises those events, sums elapsed ms / launches /
 * algorithmic FLOPs / algorithmic bytes per class into arrays of fs2_profile_classes()
 * entries, and clears the records. */
int fs2_profile_enable(fs2_handle* h, int on);
int fs2_profile_classes(void);
const char* fs2_profile_label(int cls);
int fs2_profile_read(fs2_handle* h, double* ms, int64_t* launches, double* flop, double* bytes);

/* ---- stage 1: phoneme encoder + duration predictor ------------------------------------- */
/* fastspeech.py:180-193,210: _source_mask, encoder (core/encoder.py:185-204 ->
 * attention.py:30-74, modules.py:237-248), duration_predictor (duration_predictor.py:64-86).
 *   xs [B,Tmax] i64 (0 = pad), ilens [B] i64
 *   hs [B,Tmax,adim] f32 out
 *   d_log [B,Tmax] f32 out (log-domain prediction, 0 at pads)           -- may be NULL
 *   d_int [B,Tmax] i64 out (clamp(round(exp(x)-1),0), 0 at pads)        -- may be NULL */
int fs2_encode(fs2_handle* h, const int64_t* xs, const int64_t* ilens, int B, int Tmax, float* hs, float* d_log,
               int64_t* d_int, void* ws, size_t ws_bytes, void* stream);

/* ---- stage 2: LengthRegulator (needs no handle) ---------------------------------------- */
/* core/duration_modeling/length_regulator.py:38-95 + utils/util.py:91-104.
 * Plan: per utterance, optionally scale by alpha (round half to even, :58-59), truncate to
 * ilens (:60-61), apply the all-zero -> all-one rule (:86-88; written back into `ds` when
 * mutate_ds != 0, which mirrors the reference's in-place fill_ for alpha == 1), inclusive
 * prefix sum.
 *   ds [B,Tmax] of ds_dtype; cum [B,Tmax] i32 out; olens [B] i64 out
 *   stats [2] i64 out (device): stats[0] = max_b olens[b], stats[1] = #negative durations */
int fs2_length_plan(void* ds, int ds_dtype, const int64_t* ilens, float alpha, int B, int Tmax, int mutate_ds,
                    int32_t* cum, int64_t* olens, int64_t* stats, void* stream);
/* Gather: out[b,j,:] = hs[b, min{i: cum[b,i] > j}, :] for j < olens[b], 0 for the rest of
 * [0,Lcap).  Bit-exact copy.  hs [B,Tmax,C], out [B,Lcap,C]; C % 4 == 0. */
int fs2_length_gather(const float* hs, const int32_t* cum, const int64_t* ilens, int B, int Tmax, int C, float* out,
                      int Lcap, void* stream);

/* ---- stage 3: variance adaptor + mel decoder + Postnet --------------------------------- */
/* fastspeech.py:195-238.
 *   hm [B,L,adim] f32: length-regulated encoder states (read only)
 *   olens [B] i64 or NULL.  NULL reproduces is_inference=True: decoder unmasked
 *         (fastspeech.py:221-224) and e_out / p_out unmasked.
 *   es, ps [B,L] f32 or NULL.  NULL => bucketize the predictors' own outputs (:195-196),
 *         else bucketize the given values (:200-206).
 *   before, after [B,L,odim] f32 out; e_out, p_out [B,L] f32 out (predictor values)
 *   e_ids, p_ids [B,L] i64 out (bucket indices actually embedded) -- may be NULL */
int fs2_decode(fs2_handle* h, const float* hm, const int64_t* olens, const float* es, const float* ps, int B, int L,
               float* before, float* after, float* e_out, float* p_out, int64_t* e_ids, int64_t* p_ids, void* ws,
               size_t ws_bytes, void* stream);

/* ---- stage 4: masked losses (fastspeech.py:277-333) ------------------------------------- */
/* out7 (device, f32): l1, before, after, duration, energy, pitch, total -- the order of
 * report_keys (fastspeech.py:325-333).  use_masking=True, use_weighted_masking=False.
 * scratch: >= 64 bytes of device memory (zeroed by the call). */
int fs2_masked_losses(const float* before, const float* after, const float* ys, int ld_ys_time, const float* d_out,
                      const void* ds, int ds_dtype, const float* e_out, const float* p_out, const float* es,
                      const float* ps, const int64_t* ilens, const int64_t* olens, int B, int Tmax, int L, int odim,
                      float* out7, void* scratch, void* stream);

/* ---- single operators (used by the per-kernel parity tests; same kernels as the stages) - */
/* variance_predictor.py:154-159,227-232 + fastspeech.py:218-219 */
int fs2_bucketize(const float* vals, const float* bins, int n_edges, int64_t n, int64_t* ids, void* stream);
/* F.one_hot(ids, n_bins).float(): the 4th/5th return value of _forward(is_inference=True) */
int fs2_one_hot(const int64_t* ids, int64_t n, int n_bins, float* out, void* stream);
/* out[b,t,:] = act(sum_j x[b,t+j-pad,:] . W[j] + bias) (+ resid); W [taps][N][K].
 * math_mode selects the kernel family: FS2_MATH_FP32 (CUDA cores), FS2_MATH_TF32 (kind::tf32), FS2_MATH_3XTF32 (the
 * error-compensated tensor-core family) or FS2_MATH_F16 (kind::f16); the last two run on fp16 operand planes of x and
 * w made on the fly here (inside a stage the producing kernel writes them).
 * act: 0 none, 1 relu, 2 tanh */
int fs2_op_tap_gemm(int math_mode, const float* x, int B, int L, int K, const float* w, const float* bias, int N, int taps,
                    int act, const float* resid, float* out, void* stream);
/* out = LayerNorm_N(x . w^T + bias + resid) * gamma + beta in one tcgen05 kernel (x [rows,K], w [N,K]);
 * the fused form of core/encoder.py:60-62 / :67-69: FS2_MATH_TF32 (kind::tf32 on the fp32 rows, N = 384),
 * FS2_MATH_F16 (kind::f16) and FS2_MATH_3XTF32 (error-compensated 3xF16) on operand planes made on the fly here,
 * N in {256, 384}; the plane families run as a 2-CTA cluster, each CTA normalising half of every row.
 * out_planes (nullable, plane families): the operand planes the kernel writes for the next contraction, returned
 * recombined as fp32 [rows,N] = (hi + lo) / scale (hi only in FS2_MATH_F16) */
int fs2_op_gemm_layernorm(int math_mode, const float* x, int64_t rows, int K, int N, const float* w, const float* bias,
                          const float* resid, const float* gamma, const float* beta, float eps, float* out,
                          float* out_planes, void* stream);
/* qkv [B,L,3C] (q | k | v, heads contiguous inside each) -> ctx [B,L,C]; lens NULL => no mask.  FS2_MATH_FP32: CUDA cores;
 * FS2_MATH_TF32: tcgen05 kind::tf32; FS2_MATH_F16 / FS2_MATH_3XTF32: tcgen05 kind::f16 / error-compensated 3xF16 on planes */
int fs2_op_attention(int math_mode, const float* qkv, const int64_t* lens, int B, int L, int C, int heads, float* ctx,
                     void* stream);
/* y = LayerNorm_C(x (+resid)) * g + b over the last dim (C in {256,384}) */
int fs2_op_layernorm(const float* x, const float* resid, const float* g, const float* b, float eps, int64_t rows, int C,
                     float* out, void* stream);

/* ---- train mode: dropout, BatchNorm batch statistics and the backward of every stage ------- */
/* What `model.train(); loss, _ = model(...); loss.backward()` of train_fastspeech.py:100-123 needs; fp32 on CUDA cores
 * (csrc/train.cu).  fastspeech2_b200/train.py chains these with torch.autograd.Function objects (autograd = graph plumbing
 * only).  Weights are taken in the REFERENCE's layouts (nn.Conv1d [N,K,taps], nn.Linear [N,K]); gradients are accumulated
 * (+=) into caller-zeroed buffers of the same layouts.  Activations are [B, time, channel] fp32 like the eval path. */
/* nn.Dropout (train): mask[i] = 1 keep / 0 drop from a Philox4x32-10 stream keyed by (seed, offset + i/4); out = x * mask / (1-p) */
int fs2_dropout_mask(uint8_t* mask, int64_t n, float p, uint64_t seed, uint64_t offset, void* stream);
int fs2_dropout_apply(const float* x, const uint8_t* mask, float p, float* out, int64_t n, void* stream);
/* dx = dy * f'(y) from the saved output y: act 1 = relu, 2 = tanh */
int fs2_act_backward(const float* dy, const float* y, int act, float* dx, int64_t n, void* stream);
int fs2_relu(const float* x, float* y, int64_t n, void* stream);
int fs2_add(const float* a, const float* b, float* y, int64_t n, void* stream);
int fs2_colsum(const float* x, int64_t rows, int C, float* out /* += */, void* stream);
/* Conv1d("same") / Linear: forward (core/modules.py:247-248, attention.py:48-50,74, ...), input gradient, weight + bias
 * gradient.  scratch: N*K*taps floats */
int fs2_conv_forward(const float* x, int B, int L, int K, const float* w, const float* bias, int N, int taps, int act, const float* resid,
                     float* out, float* scratch, void* stream);
int fs2_conv_dgrad(const float* dy, int B, int L, int N, const float* w, int K, int taps, float* dx, float* scratch, void* stream);
int fs2_conv_wgrad(const float* dy, const float* x, int B, int L, int N, int K, int taps, float* dw /* += */, float* dbias /* += or NULL */, void* stream);
/* nn.LayerNorm backward from the saved input rows (C in {256, 384}) */
int fs2_layernorm_backward(const float* x, const float* dy, const float* gamma, float eps, int64_t rows, int C, float* dx, float* dgamma /* += */,
                           float* dbeta /* += */, void* stream);
/* BatchNorm1d over the rows of [rows, C] with batch statistics (core/modules.py:296), optional tanh (act 2); updates the running
 * statistics in place (momentum, unbiased variance); stats [2C] = mean | biased variance; scratch: 2*C doubles */
int fs2_batchnorm_train(const float* x, int64_t rows, int C, const float* gamma, const float* beta, float eps, float momentum, int act,
                        float* running_mean, float* running_var, float* stats, float* y, void* scratch, void* stream);
int fs2_batchnorm_backward(const float* x, const float* dy, const float* stats, const float* gamma, float eps, int64_t rows, int C, float* dx,
                           float* dgamma /* += */, float* dbeta /* += */, void* scratch, void* stream);
/* strided batched fp32 GEMM over (batch, head): C = alpha * A . B, every operand as (pointer, batch stride, head stride, row
 * stride, column stride) in floats -- the attention products and their transposes in forward and backward */
int fs2_bgemm(const float* a, int64_t abs_, int64_t ahs, int64_t ars, int64_t acs, const float* b, int64_t bbs, int64_t bhs, int64_t brs, int64_t bcs,
              float* c, int64_t cbs, int64_t chs, int64_t crs, int64_t ccs, int batch, int heads, int M, int N, int K, float alpha, void* stream);
/* core/attention.py:58-69 on materialised scores [B*heads, L, L]: mask, softmax, masked_fill(0) -> p; dropout -> pd; and its backward */
int fs2_attn_softmax(const float* s, const int64_t* lens, const uint8_t* dmask, float p_drop, int B, int heads, int L, float* p, float* pd, void* stream);
int fs2_attn_softmax_backward(const float* p, const float* dpd, const uint8_t* dmask, float p_drop, int B, int heads, int L, float* ds, void* stream);
/* encoder input (fastspeech.py:65-67 + embedding.py:105-120 before its dropout) and its backward; decoder-side x + alpha*pe */
int fs2_embed_posenc(const int64_t* xs, const float* table, int n_sym, const float* pe, const float* alpha, int B, int T, int C, float* out,
                     void* stream);
int fs2_embed_backward(const int64_t* xs /* NULL: positional part only */, const float* dy, const float* pe, int B, int T, int C, int n_sym,
                       float* dtable /* += or NULL */, float* dalpha /* += */, void* stream);
int fs2_posenc_add(const float* x, const float* pe, const float* alpha, int B, int T, int C, float* y, void* stream);
/* pitch / energy embedding = Linear(n_bins -> C) on a one-hot (fastspeech.py:102,113,218-219), W [C, n_bins] */
int fs2_onehot_linear_forward(const float* x, const int64_t* ids, const float* W, const float* b, int64_t rows, int C, int n_bins, float* y,
                              void* stream);
int fs2_onehot_linear_backward(const int64_t* ids, const float* dy, int64_t rows, int C, int n_bins, float* dW /* += */, float* dbias /* += or NULL */,
                               void* stream);
/* LengthRegulator backward: dhs[b,i,:] = sum of dout[b,j,:] over the frames j expanded from phoneme i (cum from fs2_length_plan) */
int fs2_length_regulator_backward(const float* dout, const int32_t* cum, const int64_t* ilens, int B, int T, int C, int Lcap, float* dhs, void* stream);
/* predictor head Linear(C -> 1) + masked_fill(pad, 0) (duration_predictor.py:75,83-84) and its backward */
int fs2_rowdot(const float* x, const float* w, const float* bias, const int64_t* lens, int64_t rows, int L, int C, float* y, void* stream);
int fs2_rowdot_backward(const float* x, const float* w, const float* dy, const int64_t* lens, int64_t rows, int L, int C, float* dx, float* dw /* += */,
                        float* dbias /* += */, void* stream);
/* gradients of the total loss of fs2_masked_losses w.r.t. its five predicted inputs; grad_loss: device scalar dL/dloss */
int fs2_loss_backward(const float* before, const float* after, const float* ys, int ld_ys_time, const float* d_out, const void* ds, int ds_dtype,
                      const float* e_out, const float* p_out, const float* es, const float* ps, const int64_t* ilens, const int64_t* olens,
                      int B, int T, int L, int odim, const float* grad_loss, float* g_before, float* g_after, float* g_d, float* g_e, float* g_p,
                      void* stream);

/* ---- vocoder hand-off: the STFT / inverse STFT Griffin-Lim iterates ------------------------- */
/* utils/stft.py:82-151 without cuFFT, like the reference (which runs them as conv1d / conv_transpose1d with Fourier bases):
 * the two GEMMs go through fs2_op_tap_gemm, these are the kernels around them.  fastspeech2_b200/vocoder.py drives them.
 *   fs2_stft_frames       frames[b,f,k] = reflect_pad(x, n_fft/2)[b, f*hop + k]                       (:89-95)
 *   fs2_stft_magphase     spec [B*frames, ld] (real | imag) -> magnitude, phase [B, cutoff, frames]   (:105-112)
 *   fs2_istft_recombine   (magnitude, phase) -> [B*frames, ld] = mag*cos | mag*sin | 0-padding        (:115-117)
 *   fs2_istft_overlap_add overlap-add of [B, frames, n_fft] at stride hop, / window_sum where > tiny, * n_fft/hop, trimmed
 *                         by n_fft/2 at both ends -> y [B, (frames-1)*hop]                              (:119-149) */
int fs2_stft_frames(const float* x, int B, int n, int n_fft, int hop, int frames, float* out, void* stream);
int fs2_stft_magphase(const float* spec, int ld, int B, int cutoff, int frames, float* mag, float* phase, void* stream);
int fs2_istft_recombine(const float* mag, const float* phase, int B, int cutoff, int frames, int ld, float* rec, void* stream);
int fs2_istft_overlap_add(const float* frames_out, int B, int n_fft, int hop, int frames, const float* window_sum, float tiny, float* y, void* stream);

/* ---- multi-GPU exchange step: gather of the final mel shards on one rank over NVLink peer memory ----------------- */
/* Replaces what a reference user would write as torch.distributed.gather / all_gather of `after_outs` (the reference has
 * no multi-GPU path; SURVEY.md section 8e defines the step).  The root rank owns one receive buffer and exports it with
 * CUDA IPC; the other ranks map it and push their shard with a copy-engine transfer followed by a release-store of the
 * step number into a flag word; the root waits on the flags with a one-warp acquire-spin kernel.  No collective kernel
 * occupies SMs and nothing is sent to ranks that do not need it.  fastspeech2_b200/sharded.py::PeerGather drives these.
 *   fs2_peer_alloc  cudaMalloc + zero `bytes` on the current device, IPC handle (64 bytes) out
 *   fs2_peer_open   map a peer's allocation into this process (enables peer access lazily); fs2_peer_close unmaps
 *   fs2_peer_copy   asynchronous device-to-(peer-)device copy on `stream`
 *   fs2_flag_signal one-thread kernel: system-scope fence, *flag = value
 *   fs2_flag_wait   one-warp kernel: spin until flags[r] >= value for all r < n, r != skip (bounded: traps after ~10 s) */
int fs2_peer_alloc(size_t bytes, void** ptr, void* handle64);
int fs2_peer_free(void* ptr);
int fs2_peer_open(const void* handle64, void** ptr);
int fs2_peer_close(void* ptr);
int fs2_peer_copy(void* dst, const void* src, size_t bytes, void* stream);
int fs2_flag_signal(int64_t* flag, int64_t value, void* stream);
int fs2_flag_wait(const int64_t* flags, int n, int skip, int64_t value, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* FS2_B200_H_ */
