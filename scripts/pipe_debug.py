import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np, torch
import bench
import pydrex_b200 as drex
P, N = int(sys.argv[1]), 5000
dev = torch.device("cuda", 0)
kind, tables, t_edges = bench.make_workload("simple_shear", P, 10, seed=1000)
res = []
for method in (0x200, 0x400, 0x400):
    ens = drex.ParticleEnsemble(P, N, bench.PARAMS, seed=8816, device=dev)
    tol = ens.tolerances(); tol.method = method
    for k in range(int(sys.argv[2])):
        ens.update(torch.as_tensor(t_edges[0][k]).to(dev), torch.as_tensor(t_edges[1][k]).to(dev),
                   torch.as_tensor(np.ascontiguousarray(tables[k])).to(dev), kind, tol, check=False)
    torch.cuda.synchronize()
    res.append((ens.o.cpu().numpy().reshape(2 * P, 9, N), ens.f.cpu().numpy().reshape(2 * P, N), ens.stats.cpu().numpy(), ens.F.cpu().numpy()))
for a, b, name in ((res[0], res[1], "single vs pipe"), (res[1], res[2], "pipe vs pipe")):
    do = np.abs(a[0] - b[0]).max(axis=(1, 2)); df = (np.abs(a[1] - b[1]) / a[1]).max(axis=1)
    bad = np.nonzero((do > 1e-12) | (df > 1e-9))[0]
    print(name, "systems differing:", len(bad), "of", 2 * P, "first:", bad[:20])
    for s in bad[:5]:
        ng = (np.abs(a[0][s] - b[0][s]).max(axis=0) > 1e-12).sum(); nf = (np.abs(a[1][s] - b[1][s]) / a[1][s] > 1e-9).sum()
        gi = np.nonzero(np.abs(a[0][s] - b[0][s]).max(axis=0) > 1e-12)[0]
        print(f"  sys {s}: grains with different o {ng} (first {gi[:8]}), different f {nf}, do {do[s]:.2e} df {df[s]:.2e} stats a {a[2][:, s]} b {b[2][:, s]} sumf {a[1][s].sum()} {b[1][s].sum()}")
