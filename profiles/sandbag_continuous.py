import sys, numpy as np, torch
sys.path.insert(0, '/root/repo')
from pypairs_b200 import _ffi
g, n, c = 20000, 10000, 3
gen = torch.Generator(device="cuda").manual_seed(3)
lab = torch.randint(0, c, (n,), generator=gen, device="cuda")
de = torch.rand(g, generator=gen, device="cuda") < 0.02
dc = torch.randint(0, c, (g,), generator=gen, device="cuda")
x = torch.exp(torch.randn((n, g), generator=gen, device="cuda") + torch.where((lab[:, None] == dc[None, :]) & de[None, :], 1.5, 0.0)).float()
lab = lab.cpu().numpy().astype(np.int64)
sizes = np.bincount(lab, minlength=c); thr = np.ceil(sizes * 0.65).astype(np.int64); order = np.argsort(lab, kind="stable")
for _ in range(2):
    keys, cls, tm = _ffi.sandbag(x, order, sizes, np.arange(g), thr)
print("continuous cfg3 shape:", len(keys), "markers", tm)
