"""Time the attention kernels alone at a bench shape (the library's per-kernel CUDA-event profiler around the attention
launch inside fs2_op_attention; the operand preparation of the single-operator entry is outside the bracket).  Used with FS2_ATT_X2 / FS2_ATT_DEBUG to locate the bottleneck of a kernel variant."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fastspeech2_b200 import _lib

B, L, C, H = (int(v) for v in (sys.argv[1:5] if len(sys.argv) > 4 else (64, 800, 384, 2)))
lib = _lib.load()
dev = torch.device("cuda")
g = torch.Generator().manual_seed(0)
qkv = torch.randn(B, L, 3 * C, generator=g).to(dev)
ctx = torch.empty(B, L, C, device=dev)
st = _lib.stream_ptr(dev)


import ctypes as Cc
from fastspeech2_b200 import FeedForwardTransformer, synthetic_state_dict
from fastspeech2_b200.hparams import load_hp
from fastspeech2_b200.synthetic import make_batch
m = FeedForwardTransformer(68, 80, load_hp(), precision="fp32"); m.load_state_dict(synthetic_state_dict(0)); m = m.cuda().eval()
bt = make_batch(1, 8, 40, seed=1)
with torch.no_grad():
    m._forward(*[bt[k].cuda() for k in ("xs", "ilens", "olens", "ds", "es", "ps")])      # binds the profiler of this handle to the thread
h = m._handle
ncls = lib.fs2_profile_classes()
labels = [lib.fs2_profile_label(i).decode() for i in range(ncls)]


def run(mode, n=10):
    for _ in range(3):
        _lib.check(lib.fs2_op_attention(mode, _lib.ptr(qkv), None, B, L, C, H, _lib.ptr(ctx), st), "attn")
    torch.cuda.synchronize()
    lib.fs2_profile_enable(h, 1)
    for _ in range(n):
        lib.fs2_op_attention(mode, _lib.ptr(qkv), None, B, L, C, H, _lib.ptr(ctx), st)
    torch.cuda.synchronize()
    ms = (Cc.c_double * ncls)(); cnt = (Cc.c_int64 * ncls)(); fl = (Cc.c_double * ncls)(); by = (Cc.c_double * ncls)()
    lib.fs2_profile_read(h, ms, cnt, fl, by)
    lib.fs2_profile_enable(h, 0)
    i = labels.index("dec.attention")
    return ms[i] / cnt[i]


flop = 4.0 * B * L * L * C
only = [m_ for m_ in os.environ.get("PROBE_MODES", "").split(",") if m_]          # e.g. PROBE_MODES=3xf16 FS2_ATT_TRACE=1: trace that mode
for name, mode in (("f16", 3), ("3xf16", 2), ("tf32", 1)):
    if only and name not in only:
        continue
    ms = run(mode)
    print(f"X2={os.environ.get('FS2_ATT_X2','1')} DEBUG={os.environ.get('FS2_ATT_DEBUG','0')} {name}: {ms*1e3:.1f} us kernel only  ({flop/ms/1e9:.0f} TFLOP/s algorithmic)", flush=True)
