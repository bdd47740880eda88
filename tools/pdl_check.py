"""Programmatic dependent launch check: a captured step replayed many times must reproduce the eager result bit for bit
(the kernels are deterministic, so any read of not-yet-written data shows up as a difference), in every tensor-core mode,
with ragged lengths, back to back without host synchronisation in between."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from fastspeech2_b200 import FeedForwardTransformer, synthetic_state_dict  # noqa: E402
from fastspeech2_b200.hparams import load_hp  # noqa: E402
from fastspeech2_b200.synthetic import make_batch  # noqa: E402

KEYS = ("xs", "ilens", "olens", "ds", "es", "ps")
bad = 0
for prec in ("3xf16", "f16", "tf32"):
    for (B, T, L, il, ol) in ((64, 100, 800, None, None), (5, 40, 333, [40, 33, 21, 9, 3], [333, 300, 170, 64, 11])):
        m = FeedForwardTransformer(68, 80, load_hp(), precision=prec)
        m.load_state_dict(synthetic_state_dict(3))
        m = m.cuda().eval()
        bt = make_batch(B, T, L, seed=77, ilens=il, olens=ol)
        args = [bt[k].cuda() for k in KEYS]
        with torch.no_grad():
            ref = [t.clone() for t in m._forward(*args, is_inference=False)]
            g = m.graphed_forward(*args)
            outs = []
            for _ in range(30):
                o = g(*args)
                outs.append([t.clone() for t in o])
            torch.cuda.synchronize()
        n_bad = sum(0 if all(torch.equal(a, b) for a, b in zip(o, ref)) else 1 for o in outs)
        print(f"{prec} B={B} L={L}: {30 - n_bad}/30 replays bit-identical to the eager step", flush=True)
        bad += n_bad
print("PDL_CHECK", "OK" if bad == 0 else f"FAILED ({bad})")
sys.exit(0 if bad == 0 else 1)
