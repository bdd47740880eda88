"""Time the fused GEMM + LayerNorm kernel alone at a bench shape (the library's per-kernel CUDA-event profiler around the launch inside
fs2_op_gemm_layernorm; operand preparation is outside the bracket).  Used with FS2_LN_MG / FS2_LN_AMC / FS2_LN_DEBUG."""
import ctypes as Cc, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fastspeech2_b200 import _lib, FeedForwardTransformer, synthetic_state_dict
from fastspeech2_b200.hparams import load_hp
from fastspeech2_b200.synthetic import make_batch

rows, N = 51200, 384
lib = _lib.load()
m = FeedForwardTransformer(68, 80, load_hp(), precision="fp32"); m.load_state_dict(synthetic_state_dict(0)); m = m.cuda().eval()
bt = make_batch(1, 8, 40, seed=1)
with torch.no_grad():
    m._forward(*[bt[k].cuda() for k in ("xs", "ilens", "olens", "ds", "es", "ps")])      # binds the profiler of this handle to the thread
h = m._handle
ncls = lib.fs2_profile_classes()
labels = [lib.fs2_profile_label(i).decode() for i in range(ncls)]
g = torch.Generator().manual_seed(1)
for K in (384, 1024):
    x = torch.randn(rows, K, generator=g).cuda(); w = (torch.randn(N, K, generator=g) / K ** 0.5).cuda(); bias = torch.randn(N, generator=g).cuda()
    resid = torch.randn(rows, N, generator=g).cuda(); gamma = torch.ones(N).cuda(); beta = torch.zeros(N).cuda()
    out = torch.empty(rows, N, device="cuda"); planes = torch.empty(rows, N, device="cuda")
    st = _lib.stream_ptr(out.device)
    for name, mode in (("3xf16", 2), ("f16", 3)):
        def call():
            return lib.fs2_op_gemm_layernorm(mode, _lib.ptr(x), rows, K, N, _lib.ptr(w), _lib.ptr(bias), _lib.ptr(resid), _lib.ptr(gamma), _lib.ptr(beta), 1e-5,
                                             _lib.ptr(out), _lib.ptr(planes), st)
        for _ in range(3):
            _lib.check(call(), "ln")
        torch.cuda.synchronize()
        lib.fs2_profile_enable(h, 1)
        for _ in range(10):
            call()
        torch.cuda.synchronize()
        ms = (Cc.c_double * ncls)(); cnt = (Cc.c_int64 * ncls)(); fl = (Cc.c_double * ncls)(); by = (Cc.c_double * ncls)()
        lib.fs2_profile_read(h, ms, cnt, fl, by)
        lib.fs2_profile_enable(h, 0)
        i = labels.index("dec.out_proj")
        print(f"MG={os.environ.get('FS2_LN_MG','1')} AMC={os.environ.get('FS2_LN_AMC','0')} DEBUG={os.environ.get('FS2_LN_DEBUG','0')} K={K} {name}: {ms[i]/cnt[i]*1e3:.1f} us", flush=True)
