"""Per-tile error map of the fused GEMM + LayerNorm entry (debug aid): rows x N output against fp64 torch, reported per 128-row
tile and per column half."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fastspeech2_b200 import _lib

rows, K, N = (int(v) for v in (sys.argv[1:4] if len(sys.argv) > 3 else (1024, 384, 384)))
prec = sys.argv[4] if len(sys.argv) > 4 else "f16"
g = torch.Generator().manual_seed(1)
x = torch.randn(rows, K, generator=g); w = torch.randn(N, K, generator=g) / K ** 0.5; bias = torch.randn(N, generator=g)
resid = torch.randn(rows, N, generator=g) + 2.0
gamma = 1 + 0.1 * torch.randn(N, generator=g); beta = torch.randn(N, generator=g)
xc, wc, bc, rc, gc, btc = x.cuda(), w.cuda(), bias.cuda(), resid.cuda(), gamma.cuda(), beta.cuda()
y = xc.double() @ wc.double().T + bc.double() + rc.double()
want = torch.nn.functional.layer_norm(y, (N,), gc.double(), btc.double(), 1e-5).float()
out = torch.full((rows, N), float("nan"), device="cuda")
lib = _lib.load()
rcode = lib.fs2_op_gemm_layernorm(_lib.MATH_MODES[prec], _lib.ptr(xc), rows, K, N, _lib.ptr(wc), _lib.ptr(bc), _lib.ptr(rc), _lib.ptr(gc), _lib.ptr(btc),
                                  1e-5, _lib.ptr(out), None, _lib.stream_ptr(out.device))
torch.cuda.synchronize()
print("rc", rcode)
err = (out - want).abs()
for t in range((rows + 127) // 128):
    for h in range(2):
        e = err[t * 128:(t + 1) * 128, h * N // 2:(h + 1) * N // 2]
        o = out[t * 128:(t + 1) * 128, h * N // 2:(h + 1) * N // 2]
        print(f"tile {t} half {h}: nan {int(torch.isnan(o).sum())} max-err {float(torch.nan_to_num(e).max()):.3e}")
