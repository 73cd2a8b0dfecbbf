import torch
from casmtr_b200 import functional as F, pipeline
dev = torch.device('cuda:0')
wl = pipeline.Workload(832, 832, pairs=1, qt_calls=2, cas_calls=1)
host = pipeline.make_host_inputs(wl, seed=1)
call = pipeline.tree_map(lambda t: t.to(dev), host['qt'][0])
tok = [t[0].flatten(2).transpose(1, 2).contiguous() for t in (call['q'], call['k'], call['v'])]
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def run(fn, name, cold):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    F.profile_enable(True)
    for _ in range(10):
        if cold: flush.zero_()
        fn()
    torch.cuda.synchronize()
    t = F.profile_collect()
    F.profile_enable(False)
    print(name, 'cold' if cold else 'warm', {k: (round(v[0] / 10 * 1000, 1), v[1] // 10) for k, v in t.items() if v[1]})
for cold in (False, True):
    run(lambda: F.qtatt_tokens_forward(tok[0], tok[1], tok[2], (wl.h8, wl.w8), (wl.h8, wl.w8), wl.topks, wl.nh8, weight=call['weight']), 'tokens', cold)
    run(lambda: F.qtatt_forward(call['q'], call['k'], call['v'], wl.topks, wl.nh8, weight=call['weight']), 'nchw  ', cold)
