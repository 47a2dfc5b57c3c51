"""Bandwidth-bound kernels of the hot path against the HBM roofline (north star item 3), plus the
RSA-at-scale configuration (BASELINE.json configs[4]: RDM + Spearman for 1,854 x 66-D embeddings),
with the reference's NumPy/SciPy tail timed on the host cores beside it.

    python tools/bench_kernels.py [--reps 20] > profiles/rNN_kernels.json
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "vit-project_b200")]

import numpy as np  # noqa: E402
import torch  # noqa: E402

import hba  # noqa: E402
from hba import ops, rsa  # noqa: E402


class Flush:
    """L2 flush between timed launches: WRITE a 256 MB buffer (evicts everything), then READ a second one, so that
    the 126 MB the L2 is left holding are CLEAN lines - after a write-only flush every miss of the timed kernel
    first has to write a dirty flush line back, which doubles its DRAM traffic."""

    def __init__(self, dev):
        self.a = torch.zeros(64 * 1024 * 1024, device=dev)
        self.b = torch.zeros(64 * 1024 * 1024, device=dev)
        self.sink = torch.zeros((), device=dev)

    def __call__(self):
        self.a.add_(1.0)
        self.sink += self.b.sum()


def timed(fn, reps, flush):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(reps):
        flush()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))


def timed_stream(fns, replays=3):
    """Kernel throughput without the launch latency of an isolated call: the calls `fns` (the same kernel over
    DIFFERENT buffer sets, > 256 MB in total, so that every launch finds its inputs cold in the L2) are captured
    once into a CUDA graph and replayed; returns ms per launch.  This is how the kernel runs inside a step."""
    for f in fns:
        f()
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        for f in fns:
            f()
    graph.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(replays):
        graph.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (replays * len(fns))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--cpu-reps", type=int, default=5)
    a = ap.parse_args()
    print(json.dumps(measure(torch.device("cuda", 0), a.reps, a.cpu_reps), indent=1))


def measure(dev, reps=20, cpu_reps=5, with_cpu=True):
    """-> {"peak_hbm_gbs", "kernels": {name: {ms, algorithmic_bytes, achieved_gbs, frac_of_measured_peak, ...}}}.
    CUDA events around each call on the current stream, L2 flushed before every timed call, median of `reps`."""
    class _A:
        pass
    a = _A()
    a.reps, a.cpu_reps = reps, cpu_reps
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(
        os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0}
    peak = peaks["hbm_gbs"]
    flush = Flush(dev)
    out = {"peak_hbm_gbs": peak, "nominal_hbm_gbs": 8000.0, "kernels": {},
           "timing": "CUDA events around one call, median of %d, L2 flushed (256 MB written, then 256 MB read) before "
                     "each; includes the launch latency of the call (see empty_launch_us).  *_in_stream: the same kernel "
                     "replayed from a CUDA graph over rotating buffer sets (> 256 MB, inputs cold in L2), per launch - "
                     "what it costs inside a step" % reps}

    def report(name, ms, nbytes, note="", ms_stream=None):
        gbs = nbytes / (ms * 1e-3) / 1e9
        out["kernels"][name] = {"ms": ms, "algorithmic_bytes": nbytes, "achieved_gbs": gbs,
                                "frac_of_measured_peak": gbs / peak, "frac_of_nominal_8TBs": gbs / 8000.0,
                                "note": note}
        if ms_stream is not None:
            g2 = nbytes / (ms_stream * 1e-3) / 1e9
            out["kernels"][name].update(ms_in_stream=ms_stream, achieved_gbs_in_stream=g2,
                                        frac_of_measured_peak_in_stream=g2 / peak)

    # the floor of this measurement: a launch that touches 4 bytes
    tiny = torch.zeros(1, device=dev, dtype=torch.int32)
    one = torch.zeros(1, device=dev)
    out["empty_launch_us"] = 1e3 * timed(lambda: ops.nonfinite_flag(one, tiny), reps, flush)
    # ---- DoRA merge forward / backward, 1024^2 rank 32 (NEW:447-463)
    g = torch.Generator(device=dev).manual_seed(0)
    n, r = 1024, 32
    D = torch.randn(n, n, device=dev, generator=g)
    D = D / D.norm(dim=0)
    A = torch.randn(r, n, device=dev, generator=g) * 0.03
    Bm = torch.randn(n, r, device=dev, generator=g) * 0.1
    m = torch.rand(n, device=dev, generator=g) + 0.5
    w_t = torch.empty(n, n, device=dev)
    w, wt = ops.Operand.empty(n, n, False, dev), ops.Operand.empty(n, n, False, dev)
    ms = timed(lambda: ops.dora_merge_fwd(D, A, Bm, m, 0.5, 1e-8, w_t_f32=w_t), a.reps, flush)
    report("dora_merge_fwd_1024_fp32_only", ms, 4 * n * n * 2 + 4 * (2 * r * n + n),
           "read D + write Wt (SURVEY 8d: 8.66 MB)")
    ms = timed(lambda: ops.dora_merge_fwd(D, A, Bm, m, 0.5, 1e-8, w_t_f32=w_t, w=w, wt=wt), a.reps, flush)
    sets = [(D.clone(), torch.empty(n, n, device=dev), ops.Operand.empty(n, n, False, dev),
             ops.Operand.empty(n, n, False, dev), torch.randn(n, n, device=dev, generator=g),
             torch.empty(n, n, device=dev)) for _ in range(14)]   # 14 x 20 MB
    ms_s = timed_stream([(lambda t=t: ops.dora_merge_fwd(t[0], A, Bm, m, 0.5, 1e-8, w_t_f32=t[1], w=t[2], wt=t[3]))
                         for t in sets])
    report("dora_merge_fwd_1024_with_bf16_operands", ms, 4 * n * n * 2 + 2 * 2 * n * n + 4 * (2 * r * n + n),
           "+ W and W^T bf16 GEMM operands written by the same kernel", ms_stream=ms_s)
    G = torch.randn(n, n, device=dev, generator=g)
    dm, dA, dB, ws = torch.empty_like(m), torch.empty_like(A), torch.empty_like(Bm), torch.empty(n, n, device=dev)
    ms = timed(lambda: ops.dora_merge_bwd(G, D, A, Bm, m, 0.5, 1e-8, dm, dA, dB, ws), a.reps, flush)
    ms_s = timed_stream([(lambda t=t: ops.dora_merge_bwd(t[4], t[0], A, Bm, m, 0.5, 1e-8, dm, dA, dB, t[5]))
                         for t in sets])
    del sets
    report("dora_merge_bwd_1024", ms, 4 * n * n * 2 + 4 * (4 * r * n + 2 * n),
           "algorithmic: read G and D once, write dm / dA / dB (the [out/32][in][r] partial sums of dB make one more "
           "4 MB round trip through the L2); 2 launches", ms_stream=ms_s)
    # ---- LayerNorm forward 8224 x 1024 (fp32 in, bf16 out)
    x = torch.randn(8224, 1024, device=dev, generator=g)
    gam, bet = torch.ones(1024, device=dev), torch.zeros(1024, device=dev)
    y = ops.Operand.empty(8224, 1024, False, dev)
    ms = timed(lambda: ops.layernorm_fwd(x, 8224, 1024, gam, bet, 1e-5, y=y), a.reps, flush)
    sets = [(torch.randn(8224, 1024, device=dev, generator=g), ops.Operand.empty(8224, 1024, False, dev))
            for _ in range(6)]   # 6 x 50 MB
    ms_s = timed_stream([(lambda t=t: ops.layernorm_fwd(t[0], 8224, 1024, gam, bet, 1e-5, y=t[1])) for t in sets])
    del sets
    report("layernorm_fwd_8224x1024", ms, 8224 * 1024 * 6, "fp32 read + bf16 write", ms_stream=ms_s)
    # ---- cosine + MSE head, B = 32 (0.32 MB: latency bound)
    img, txt = torch.randn(32, 768, device=dev, generator=g), torch.randn(66, 768, device=dev, generator=g)
    ls = torch.tensor([4.6052], device=dev)
    pred, tgt, loss = torch.empty(32, 66, device=dev), torch.randn(32, 66, device=dev, generator=g), torch.empty(1, device=dev)
    ms = timed(lambda: ops.cos_head_fwd(img, txt, ls, pred, tgt, loss), a.reps, flush)
    report("cos_mse_head_fwd_B32", ms, 4 * (32 * 768 + 66 * 768 + 2 * 32 * 66), "latency bound (2 launches)")
    # ---- AdamW over the 9 DoRA tensors (183,040 parameters)
    shapes = [(1024,), (32, 1024), (1024, 32)] * 2 + [(768,), (32, 768), (768, 32)]
    ps = [torch.randn(*s, device=dev, generator=g) for s in shapes]
    gs, m1, m2 = [torch.randn_like(p) for p in ps], [torch.zeros_like(p) for p in ps], [torch.zeros_like(p) for p in ps]
    flat = [t.data_ptr() for i in range(9) for t in (ps[i], gs[i], m1[i], m2[i])]
    table = torch.tensor(flat, dtype=torch.int64, device=dev)
    sizes = torch.tensor([p.numel() for p in ps], dtype=torch.int64, device=dev)
    total = int(sizes.sum())
    ms = timed(lambda: ops.adamw_multi(table, sizes, 9, total, 3e-4, 0.9, 0.999, 1e-8, 0.01, 1), a.reps, flush)
    report("adamw_multi_183040", ms, total * 4 * 7, "latency bound (one launch for 9 tensors)")
    # ---- RSA at scale: N = 1854, D = 66 -> P = 1,717,731 pairs
    for N in (48, 1854):
        rng = np.random.default_rng(N)
        E = rng.standard_normal((N, 66)).astype(np.float32)
        ref = 1 - np.corrcoef(rng.standard_normal((N, 66)) + 0.5 * E)
        np.fill_diagonal(ref, 0)
        ev = rsa.RSAEvaluator(ref, dev)
        Ed = torch.from_numpy(E).to(dev)
        P = N * (N - 1) // 2
        ms = timed(lambda: ev.rho_device(Ed, want_rdm=False), a.reps, flush)
        report(f"rsa_rdm_rank_spearman_N{N}", ms, 4 * 66 * N + 40 * P,
               "RDM (f64) + average-tie ranking + Pearson on ranks; bytes = 4*66*N + 40*P (SURVEY 8d)")
        rho = float(ev.rho.cpu())
        if not with_cpu:
            out["kernels"][f"rsa_rdm_rank_spearman_N{N}"].update(rho_gpu=rho, checkpoints_per_s=1e3 / ms)
            continue
        from scipy.stats import spearmanr
        t0 = time.perf_counter()
        for _ in range(a.cpu_reps):
            model_rdm = 1 - np.corrcoef(E)
            np.fill_diagonal(model_rdm, 0)
            iu = np.triu_indices_from(ref, k=1)
            rho_cpu, _ = spearmanr(ref[iu], model_rdm[iu])
        cpu_ms = (time.perf_counter() - t0) / a.cpu_reps * 1e3
        k = out["kernels"][f"rsa_rdm_rank_spearman_N{N}"]
        k.update(cpu_reference_ms=cpu_ms, cpu_cores=os.cpu_count(), speedup_vs_cpu=cpu_ms / ms,
                 rho_gpu=rho, rho_cpu=float(rho_cpu), rho_abs_diff=abs(rho - float(rho_cpu)),
                 checkpoints_per_s=1e3 / ms)
    return out


if __name__ == "__main__":
    main()
