"""RSA evaluation on the GPU (replaces the NumPy/SciPy tail of the reference's behavioral_RSA,
NEW:625-652, and of compute_rsa_score, MEAS:298-355):

    model_rdm = 1 - corrcoef(E)  (float64, diag 0)  ->  upper triangle (k=1, row-major)
    rho = Pearson(rank_avg(reference_tri), rank_avg(model_tri))     (= scipy.stats.spearmanr)

Only the Student-t survival function for the p-value is evaluated on the host (one scalar).
"""
from __future__ import annotations

import numpy as np
import torch

from . import ops


class RSAEvaluator:
    """Holds the ranked reference RDM and the workspaces for repeated evaluations of N x D
    embeddings against one reference RDM."""

    def __init__(self, reference_rdm, device="cuda"):
        ref = np.asarray(reference_rdm, dtype=np.float64)
        if ref.ndim != 2 or ref.shape[0] != ref.shape[1]:
            raise ValueError("reference_rdm must be a square matrix")
        self.N = ref.shape[0]
        self.P = self.N * (self.N - 1) // 2
        self.device = torch.device(device)
        iu = np.triu_indices(self.N, k=1)
        ref_tri = torch.from_numpy(np.ascontiguousarray(ref[iu])).to(self.device)
        self.rank_ws = torch.empty(max(ops.rank_workspace_bytes(self.P), 256), dtype=torch.uint8,
                                   device=self.device)
        self.ref_ranks = torch.empty(self.P, dtype=torch.float64, device=self.device)
        ops.rank_avg_f64(ref_tri, self.ref_ranks, self.rank_ws)
        self.tri = torch.empty(self.P, dtype=torch.float64, device=self.device)
        self.ranks = torch.empty(self.P, dtype=torch.float64, device=self.device)
        self.rdm = torch.empty(self.N, self.N, dtype=torch.float64, device=self.device)
        self.pearson_ws = torch.empty(5 * 1024, dtype=torch.float64, device=self.device)
        self.rho = torch.empty(1, dtype=torch.float64, device=self.device)

    def rho_device(self, emb, want_rdm=True):
        """Launches the whole chain on the current stream; returns the device scalar rho."""
        if emb.shape[0] != self.N:
            raise ValueError(f"expected {self.N} embeddings, got {emb.shape[0]}")
        emb = emb.detach().to(self.device, torch.float32).contiguous()
        if self.P > 2048:
            # RSA at scale: one fused chain (keys straight from the RDM kernel, ranks consumed in sorted order)
            ops.rdm_spearman(emb, self.ref_ranks, self.rho, self.rank_ws, rdm=self.rdm if want_rdm else None)
            return self.rho
        ops.rdm_f64(emb, self.rdm if want_rdm else None, self.tri)
        ops.rank_avg_f64(self.tri, self.ranks, self.rank_ws)
        ops.pearson_f64(self.ref_ranks, self.ranks, self.rho, self.pearson_ws)
        return self.rho

    def __call__(self, emb, want_rdm=True):
        """-> (rho, p_value, model_rdm ndarray | None), like behavioral_RSA's return (NEW:654)."""
        with torch.cuda.device(self.device):
            rho = float(self.rho_device(emb, want_rdm).cpu())
        rdm = self.rdm.cpu().numpy() if want_rdm else None
        return rho, spearman_pvalue(rho, self.P), rdm


def spearman_pvalue(rho, n):
    """Two-sided p-value of scipy.stats.spearmanr: t = r sqrt(dof / ((r+1)(1-r))), dof = n-2."""
    from scipy import special
    dof = n - 2
    if dof <= 0 or rho != rho:   # scipy returns (nan, nan) for constant / NaN input
        return float("nan")
    denom = (rho + 1.0) * (1.0 - rho)
    with np.errstate(divide="ignore", invalid="ignore"):
        t = rho * np.sqrt(max(dof / denom, 0.0)) if denom > 0 else np.inf * np.sign(rho)
    return float(2.0 * special.stdtr(dof, -abs(t)))


def rsa_from_embeddings(emb, reference_rdm):
    return RSAEvaluator(reference_rdm, emb.device)(emb)
