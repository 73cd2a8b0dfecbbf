import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from casmtr_b200 import functional as F, synth
dev = torch.device('cuda:0')
for B in (1, 2, 16):
    qs, ks, vs, wt = synth.qtatt_inputs(B, 256, 104, 104, 3, seed=3)
    args = ([t.to(dev) for t in qs], [t.to(dev) for t in ks], [t.to(dev) for t in vs], [32, 16, 8], 8)
    for _ in range(3):
        F.qtatt_forward(*args, weight=wt.to(dev))
    torch.cuda.synchronize()
