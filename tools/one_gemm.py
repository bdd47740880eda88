import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fastspeech2_b200 import _lib
lib = _lib.load(); dev = torch.device("cuda")
B, L, K, N, taps = 64, 800, int(sys.argv[1]), int(sys.argv[2]), 1
x = torch.randn(B, L, K, device=dev); w = torch.randn(taps, N, K, device=dev) * 0.05; bias = torch.randn(N, device=dev)
out = torch.empty(B, L, N, device=dev); st = _lib.stream_ptr(dev)
for _ in range(3):
    lib.fs2_op_tap_gemm(1, _lib.ptr(x), B, L, K, _lib.ptr(w), _lib.ptr(bias), N, taps, 0, None, _lib.ptr(out), st)
torch.cuda.synchronize()
