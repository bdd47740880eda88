"""One forward step of the bench workload bracketed by cudaProfilerStart/Stop, for
`ncu --profile-from-start off ...` (launch list or a full capture of one kernel)."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from bench import WORKLOADS  # noqa: E402
from fastspeech2_b200 import FeedForwardTransformer, synthetic_state_dict  # noqa: E402
from fastspeech2_b200.hparams import load_hp  # noqa: E402
from fastspeech2_b200.synthetic import make_batch  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="c2")
ap.add_argument("--precision", default="tf32")
ap.add_argument("--warmup", type=int, default=2)
a = ap.parse_args()
B, T, L = WORKLOADS[a.workload]
m = FeedForwardTransformer(68, 80, load_hp(), precision=a.precision)
m.load_state_dict(synthetic_state_dict(0))
m = m.cuda().eval()
bt = make_batch(B, T, L, seed=1234)
inp = [bt[k].cuda() for k in ("xs", "ilens", "olens", "ds", "es", "ps")]
with torch.no_grad():
    for _ in range(a.warmup):
        m._forward(*inp)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStart()
    m._forward(*inp)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()
print("profiled one step of", a.workload, a.precision)
