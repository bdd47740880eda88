"""One step that touches the remaining kernel families, bracketed for `ncu --profile-from-start off`:
tf32 c2 forward() (row_norm, variance_embed_add, embed_posenc, length_plan/gather, attention_fp32 in the encoder,
masked-loss kernels), to_one_hot (bucketize, one_hot) and -- with --fp32 -- an fp32-mode forward (gemm_fp32)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fastspeech2_b200 import FeedForwardTransformer, synthetic_state_dict
from fastspeech2_b200.hparams import load_hp
from fastspeech2_b200.synthetic import make_batch
prec = "fp32" if "--fp32" in sys.argv else "tf32"
m = FeedForwardTransformer(68, 80, load_hp(), precision=prec); m.load_state_dict(synthetic_state_dict(0)); m = m.cuda().eval()
bt = make_batch(64, 100, 800, seed=1234)
a = [bt[k].cuda() for k in ("xs", "ilens", "ys", "olens", "ds", "es", "ps")]
with torch.no_grad():
    m(*a); torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStart()
    m(*a)
    m.energy_predictor.to_one_hot(a[5])
    torch.cuda.synchronize(); torch.cuda.cudart().cudaProfilerStop()
print("ok", prec)
