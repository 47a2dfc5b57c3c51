import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "vit-project_b200")]
import numpy as np, torch
from hba import ops, rsa
DEV = torch.device("cuda", 0)
N = 300
rng = np.random.default_rng(N)
E = rng.standard_normal((N, 66)).astype(np.float32)
E[3, 5] = np.nan
ref_rdm = 1 - np.corrcoef(rng.standard_normal((N, 66)))
np.fill_diagonal(ref_rdm, 0)
ev = rsa.RSAEvaluator(ref_rdm, DEV)
P = N * (N - 1) // 2
ranks = torch.zeros(P, dtype=torch.float64, device=DEV)
emb = torch.from_numpy(E).to(DEV)
ops.rdm_spearman(emb, ev.ref_ranks, ev.rho, ev.rank_ws, rdm=ev.rdm, ranks=ranks)
torch.cuda.synchronize()
print("rho", float(ev.rho), "rdm nan", int(torch.isnan(ev.rdm).sum()), "ranks nan", int(torch.isnan(ranks).sum()),
      "ranks max", float(torch.nan_to_num(ranks).max()))
ws = ev.rank_ws
# keys live at the start of the workspace
keys = ws[: P * 8].view(torch.int64)
print("k0 top keys", [hex(int(k) & (2**64 - 1)) for k in keys.sort().values[-3:]], [hex(int(k) & (2**64-1)) for k in keys.sort().values[:3]])
tri = torch.empty(P, dtype=torch.float64, device=DEV)
ops.rdm_f64(emb, None, tri)
print("tri nan", int(torch.isnan(tri).sum()), hex(int(tri[torch.isnan(tri)][:1].view(torch.int64)) & (2**64-1)) if torch.isnan(tri).any() else None)
r2 = torch.empty(P, dtype=torch.float64, device=DEV)
ops.rank_avg_f64(tri, r2)
print("rank_avg nan", int(torch.isnan(r2).sum()))
