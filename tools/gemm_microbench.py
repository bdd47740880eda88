"""Time fs2_op_tap_gemm for a list of shapes (CUDA events, L2-cold-ish: big operands)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fastspeech2_b200 import _lib
lib = _lib.load()
dev = torch.device("cuda")
shapes = [  # (B, L, K, N, taps, resid, mode)
    (64, 800, 32, 1152, 1, 0, 1), (64, 800, 96, 1152, 1, 0, 1), (64, 800, 384, 1152, 1, 0, 1), (64, 800, 1024, 1152, 1, 0, 1),
    (64, 800, 32, 384, 1, 0, 1), (64, 800, 384, 384, 1, 0, 1), (64, 800, 384, 384, 1, 1, 1), (64, 800, 1024, 384, 1, 1, 1),
    (64, 800, 32, 1024, 1, 0, 1), (64, 800, 384, 1024, 1, 0, 1), (64, 800, 384, 1024, 9, 0, 1),
    (64, 800, 256, 256, 3, 0, 2), (64, 100, 256, 1024, 9, 0, 2), (64, 100, 256, 768, 1, 0, 2),
]
for (B, L, K, N, taps, resid, mode) in shapes:
    x = torch.randn(B, L, K, device=dev); w = torch.randn(taps, N, K, device=dev) * 0.05; bias = torch.randn(N, device=dev)
    r = torch.randn(B, L, N, device=dev) if resid else None
    out = torch.empty(B, L, N, device=dev)
    st = _lib.stream_ptr(dev)
    def run():
        _lib.check(lib.fs2_op_tap_gemm(mode, _lib.ptr(x), B, L, K, _lib.ptr(w), _lib.ptr(bias), N, taps, 0, _lib.ptr(r), _lib.ptr(out), st), "gemm")
    for _ in range(3): run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): run()
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 100
    flop = 2.0 * B * L * K * N * taps
    byts = 4.0 * (B * L * K + taps * N * K + B * L * N * (2 if resid else 1))
    print(f"M={B*L:6d} K={K:5d} N={N:5d} taps={taps} resid={resid} mode={mode}: {us:8.1f} us  {flop/us/1e6:7.1f} TF  {byts/us/1e3:7.0f} GB/s")
