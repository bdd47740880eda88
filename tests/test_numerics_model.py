"""CPU model of the operand formats the tensor-core families use (no GPU): what the fp16 hi / lo split of the
error-compensated "3xF16" family and the plain fp16 copies of the "f16" family can and cannot represent.  These are
properties of the number formats, so they are checked with numpy; the kernels themselves are checked on the GPU
(tests/test_gpu_parity.py)."""
import numpy as np


def split_f16(x):
    """hi = rn_fp16(clamp(x)), lo = rn_fp16(x - hi): split_rows_f16_kernel / split_f16_kernel (gemm_tc.cu)."""
    xc = np.clip(x.astype(np.float32), -65504.0, 65504.0)
    hi = xc.astype(np.float16)
    lo = (xc - hi.astype(np.float32)).astype(np.float16)
    return hi, lo


def dot_3xf16(x, w):
    """The three products the kernel accumulates (lo*hi + hi*lo + hi*hi), summed exactly (float64) so that only the
    operand representation error remains (the tensor core's own fp32 accumulation is a separate, measured term)."""
    xh, xl = (a.astype(np.float64) for a in split_f16(x))
    wh, wl = (a.astype(np.float64) for a in split_f16(w))
    return xl @ wh + xh @ wl + xh @ wh


def test_split_carries_22_bits_above_the_fp16_subnormal_range():
    g = np.random.RandomState(0)
    x = g.randn(64, 2304).astype(np.float32) * 2            # LayerNorm-scale activations
    hi, lo = split_f16(x)
    err = np.abs(hi.astype(np.float64) + lo.astype(np.float64) - x)
    big = np.abs(x) >= 2.0 ** -3                              # lo = O(2^-12 x) is a normal fp16 there
    assert (err[big] / np.abs(x[big])).max() <= 2.0 ** -21   # 22 significant bits
    assert err[~big].max() <= 2.0 ** -25                      # below: lo is subnormal, absolute error <= half its spacing (2^-24)


def test_3xf16_dot_error_floor():
    """Weights of a trained layer (|w| < 0.06) have a *subnormal* fp16 lo part, so each carries an absolute error of up
    to 2^-25 = 3e-8 instead of a relative 2^-22.  The dot product therefore has an absolute error floor of about
    sqrt(K) * rms(x) * 2e-8 (2e-6 at K = 2304, rms 2), independent of the weight scale: fp32-class (1e-6 relative) for
    ordinary layers with results of O(1), an order of magnitude below the tensor core's own accumulation error
    (2e-5..2e-4, tests/test_gpu_parity.py), and degrading towards the 10-bit class only for layers whose weights are
    all below ~1e-3.  (DESIGN.md section 4 lists the power-of-two operand pre-scaling that removes the floor.)"""
    g = np.random.RandomState(1)
    K = 2304
    x = g.randn(128, K).astype(np.float32) * 2
    errs = {}
    for wscale in (1.0 / np.sqrt(K), 1e-3, 1e-4):
        w = (g.randn(K, 96) * wscale).astype(np.float32)
        exact = x.astype(np.float64) @ w.astype(np.float64)
        errs[wscale] = np.abs(dot_3xf16(x, w) - exact).max()
        assert errs[wscale] <= 1e-5, (wscale, errs[wscale])
    # ordinary layer: result rms = 2, so the floor is ~3e-6 relative; plain 10-bit operands are ~300x coarser
    w = (g.randn(K, 96) / np.sqrt(K)).astype(np.float32)
    exact = x.astype(np.float64) @ w.astype(np.float64)
    plain = x.astype(np.float16).astype(np.float64) @ w.astype(np.float16).astype(np.float64)
    assert np.abs(plain - exact).max() > 100 * np.abs(dot_3xf16(x, w) - exact).max()


def test_fp16_copy_is_at_least_as_fine_as_tf32_truncation():
    """The f16 family reads round-to-nearest fp16 copies; kind::tf32 reads the top 19 bits of the fp32 word (truncation):
    same 10-bit mantissa, half the worst-case error -- which is why the f16 mode shares the tf32 mode's tolerance."""
    g = np.random.RandomState(2)
    x = (g.randn(1 << 16) * 3).astype(np.float32)
    e16 = np.abs(x.astype(np.float16).astype(np.float64) - x)
    tf32 = (x.view(np.uint32) & np.uint32(0xFFFFE000)).view(np.float32)
    e32 = np.abs(tf32.astype(np.float64) - x)
    assert e16.max() <= e32.max() and e16.mean() < 0.6 * e32.mean()
    # range: the clamp only matters beyond +-65504; LayerNorm outputs and ReLU(conv) activations are O(1..100)
    assert np.isfinite(np.clip(np.float32(1e6), -65504, 65504).astype(np.float16))
