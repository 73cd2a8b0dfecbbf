python - <<'PY'
import torch, time
from casmtr_b200 import functional as F
dev=torch.device('cuda')
g=torch.Generator().manual_seed(0)
f0=torch.randn(1,10816,256,generator=g).to(dev); f1=torch.randn(1,10816,256,generator=g).to(dev)
for _ in range(3): F.coarse_match_forward(f0,f1,0.1)
torch.cuda.synchronize()
F.profile_collect(); F.profile_enable(True)
e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): F.coarse_match_forward(f0,f1,0.1)
e1.record(); torch.cuda.synchronize(); F.profile_enable(False)
print('coarse_match ms per call', e0.elapsed_time(e1)/10, F.profile_collect()['coarse_match'])
# torch reference on GPU
def ref():
    sim=torch.einsum('nlc,nsc->nls', f0/16, f1/16)/0.1
    a=torch.softmax(sim,1); b=torch.softmax(sim,2)
    conf=a*b
    return b.max(dim=2), a.max(dim=1), conf
for _ in range(2): ref()
torch.cuda.synchronize(); e0.record()
for _ in range(5): ref()
e1.record(); torch.cuda.synchronize()
print('torch reference formulation ms per call', e0.elapsed_time(e1)/5)
PY
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'coarse_rowstats' -c 1 -f -o gpurun_out/r01t_coarse_match python -c "
import torch
from casmtr_b200 import functional as F
dev=torch.device('cuda')
g=torch.Generator().manual_seed(0)
f0=torch.randn(1,10816,256,generator=g).to(dev); f1=torch.randn(1,10816,256,generator=g).to(dev)
F.coarse_match_forward(f0,f1,0.1); torch.cuda.synchronize()
" > gpurun_out/r01t_ncu.log 2>&1
ncu -i gpurun_out/r01t_coarse_match.ncu-rep --page raw --csv > gpurun_out/r01t_coarse_match_raw.csv 2>/dev/null
