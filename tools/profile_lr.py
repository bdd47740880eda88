"""LengthRegulator at BASELINE config 5 bracketed by cudaProfilerStart/Stop for ncu."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fastspeech2_b200 import LengthRegulator
g = torch.Generator().manual_seed(1234)
hs = torch.randn(256, 100, 256, generator=g).cuda(); ds = torch.randint(1, 16, (256, 100), generator=g).cuda()
il = torch.full((256,), 100, dtype=torch.int64).cuda()
lr = LengthRegulator()
for _ in range(2): lr(hs, ds, il, alpha=4.0)
torch.cuda.synchronize(); torch.cuda.cudart().cudaProfilerStart()
out = lr(hs, ds, il, alpha=4.0)
torch.cuda.synchronize(); torch.cuda.cudart().cudaProfilerStop()
print("Lmax", out.shape[1])
