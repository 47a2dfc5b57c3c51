"""CPU tests of the multi-process host logic: the sweep scheduler (hba.sweep: condition lists of the
reference drivers, LPT order, per-condition config / directory layout, a 2-worker pool with a failing
condition), the data-parallel gradient buckets on a world-size-2 `gloo` group (hba.dp), and bench.py's
reference arm under torchrun with 2 ranks."""
import json
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.timeout(900)      # worker pools / spawned ranks / torchrun: a hang must fail, not stall the suite


# ------------------------------------------------------------------------------- sweep scheduler
def test_condition_lists_match_the_reference_drivers():
    from hba import sweep
    single = sweep.single_epoch_conditions()
    assert [c["training_run"] for c in single] == list(range(1, 99)) and all(c["perturb_length"] == 1 for c in single)
    grid = sweep.length_grid_conditions()
    assert len(grid) == 136 and len({(c["training_run"], c["perturb_length"]) for c in grid}) == 136
    starts = {c["training_run"] for c in grid}
    assert starts == {1, 2, 3, 6, 7, 8, 10, 20, 30, 40, 50, 60, 70, 80, 90, 13, 16, 19, 58, 94, 22}
    assert {c["perturb_length"] for c in grid if c["training_run"] == 13} == {5, 10, 20, 30, 40, 50}
    assert {c["perturb_length"] for c in grid if c["training_run"] == 1} == {2, 5, 10, 20, 30, 40, 50}
    assert [c["perturb_length"] for c in grid if c["training_run"] == 22] == [5]


def test_lpt_plan_is_balanced_and_complete():
    from hba import sweep
    grid = sweep.length_grid_conditions()
    plan, loads = sweep.lpt_assign(grid, 8)
    flat = [(c["training_run"], c["perturb_length"]) for p in plan for c in p]
    assert sorted(flat) == sorted((c["training_run"], c["perturb_length"]) for c in grid)
    ideal = sum(loads) / 8
    assert max(loads) <= ideal * 1.02          # LPT on 136 jobs: within 2 % of the ideal makespan
    order = sweep.lpt_order(grid)
    costs = [sweep.expected_epochs(c) for c in order]
    assert costs == sorted(costs, reverse=True)


def test_condition_config_layouts(tmp_path):
    from hba import sweep
    base = {"output_base_directory": str(tmp_path), "perturb_type": "label_shuffle", "perturb_length": 1, "cuda": 1}
    cfg = sweep.condition_config(base, {"training_run": 15, "perturb_length": 1}, "sweep")
    d = os.path.join(str(tmp_path), "training_run15")              # SWEEP:198-207
    assert cfg["training_res_path"] == os.path.join(d, "training_res_run15.csv")
    assert cfg["dora_parameters_path"] == os.path.join(d, "dora_params_run15")
    assert cfg["random_state_path"] == os.path.join(d, "random_states_run15")
    assert cfg["checkpoint_path"] == os.path.join(d, "model_checkpoint_run15.pth")
    assert cfg["resume_from_epoch"] == 14 and cfg["training_run"] == 15 and os.path.isdir(d)
    cfg = sweep.condition_config(base, {"training_run": 1, "perturb_length": 20}, "length")
    d = os.path.join(str(tmp_path), "label_shuffle_e1_l20")        # LEN:128-137
    assert cfg["training_res_path"] == os.path.join(d, "training_res.csv")
    assert cfg["dora_parameters_path"] == os.path.join(d, "dora_params_1") and cfg["resume_from_epoch"] == 0
    assert cfg["perturb_length"] == 20 and "training_run" not in base


def _fake_condition(cfg):
    """Stands in for run_behavioral_training in the pool test (module level: picklable under spawn)."""
    if cfg["training_run"] == 3:
        raise RuntimeError("synthetic failure of run 3")
    with open(cfg["training_res_path"], "w") as f:
        f.write(f"{cfg['training_run']},{cfg['perturb_length']},{cfg['resume_from_epoch']},{os.getpid()}\n")


def test_two_worker_pool_runs_every_condition_and_survives_a_failure(tmp_path):
    from hba import sweep
    conds = sweep.single_epoch_conditions(1, 6)
    base = {"output_base_directory": str(tmp_path), "perturb_type": "random_target", "cuda": 0}
    logs = []
    res = sweep.run_sweep(base, conds, [None, None], run_fn=_fake_condition, log=logs.append)
    assert [r["condition"]["training_run"] for r in res] == [1, 2, 3, 4, 5, 6]
    assert [r["ok"] for r in res] == [True, True, False, True, True, True]
    assert "synthetic failure" in res[2]["error"]
    pids = set()
    for e in (1, 2, 4, 5, 6):
        row = open(os.path.join(str(tmp_path), f"training_run{e}", f"training_res_run{e}.csv")).read().strip().split(",")
        assert row[:3] == [str(e), "1", str(e - 1)]
        pids.add(row[3])
    assert len(pids) <= 2 and os.getpid() not in {int(p) for p in pids}
    assert {r["worker"] for r in res} <= {0, 1}
    assert "5 successful, 1 failed" in logs[-1]


# ------------------------------------------------------------------------------- gloo, world size 2
def _report_threads(cfg):
    import json
    import torch
    with open(os.path.join(cfg["output_base_directory"], f"threads_{cfg['training_run']}.json"), "w") as f:
        json.dump({"torch": torch.get_num_threads(), "omp": os.environ.get("OMP_NUM_THREADS"),
                   "wait": os.environ.get("OMP_WAIT_POLICY")}, f)


def test_workers_start_with_capped_cpu_thread_pools(tmp_path, monkeypatch):
    """Eight workers starting at once, each with OpenMP / MKL pools sized for all host cores, stalled 700 s in the
    image pipeline of the full-grid run on 8 B200s (profiles/r02_grid_full_136_n8.json; reproduced on the CPU: 400
    images decode in 2 s alone, 145 s in each of 8 concurrent default processes, 2.7 s with the caps).  run_sweep
    hands every worker cores / workers threads (at most 4) through the environment of the spawn and restores its own."""
    import json
    from hba import sweep
    monkeypatch.setenv("OMP_NUM_THREADS", "7")
    monkeypatch.delenv("OMP_WAIT_POLICY", raising=False)
    conds = [{"training_run": e, "perturb_length": 1} for e in (1, 2)]
    res = sweep.run_sweep({"output_base_directory": str(tmp_path)}, conds, [None, None], run_fn=_report_threads,
                          log=lambda *_: None)
    assert all(r["ok"] for r in res), [r["error"] for r in res]
    want = sweep.cpu_threads_per_worker(2)
    assert 1 <= want <= 4 and sweep.cpu_threads_per_worker(10 ** 6) == 1
    for e in (1, 2):
        got = json.load(open(os.path.join(str(tmp_path), f"threads_{e}.json")))
        assert got == {"torch": want, "omp": str(want), "wait": "passive"}
    assert os.environ["OMP_NUM_THREADS"] == "7" and "OMP_WAIT_POLICY" not in os.environ   # the parent's own, restored


def _report_visible_devices(cfg):
    with open(os.path.join(cfg["output_base_directory"], f"visible_{cfg['training_run']}.txt"), "w") as f:
        f.write(f"{os.environ.get('CUDA_VISIBLE_DEVICES')}|{cfg['cuda']}")
    import time
    time.sleep(1.0)    # (both workers get a condition)


def test_workers_are_pinned_through_the_launchers_own_device_list(tmp_path, monkeypatch):
    """`devices` of run_sweep are logical indices of the CALLING process: under a launcher that already restricted
    the job (CUDA_VISIBLE_DEVICES="4,5", UUIDs, MIG slices) worker i gets the i-th entry of that list - never a
    physical GPU outside it - and maps config['cuda'] onto the one device it then sees."""
    from hba import sweep
    assert sweep.pinned_device_env(1, "4,5") == "5" and sweep.pinned_device_env(0, " GPU-aa , GPU-bb ") == "GPU-aa"
    assert sweep.pinned_device_env(0, "MIG-GPU-1/3/0") == "MIG-GPU-1/3/0"
    for bad, cur in ((2, "4,5"), (0, ""), (-1, "0,1")):
        with pytest.raises(ValueError):
            sweep.pinned_device_env(bad, cur)
    monkeypatch.setenv("CUDA_VISIBLE_DEVICES", "6,3")
    conds = [{"training_run": e, "perturb_length": 1} for e in (1, 2, 3, 4)]
    res = sweep.run_sweep({"output_base_directory": str(tmp_path), "cuda": 1}, conds, [0, 1],
                          run_fn=_report_visible_devices, log=lambda *_: None)
    assert all(r["ok"] for r in res), [r["error"] for r in res]
    for r in res:      # whichever worker took a condition: it saw exactly its own entry of the launcher's list
        seen = open(os.path.join(str(tmp_path), f"visible_{r['condition']['training_run']}.txt")).read()
        assert seen == {0: "6|0", 1: "3|0"}[r["worker"]], (r["worker"], seen)
    assert {r["worker"] for r in res} == {0, 1}                  # (four 1-second conditions: both workers get work)
    assert os.environ["CUDA_VISIBLE_DEVICES"] == "6,3"           # the parent's own list is untouched
    with pytest.raises(ValueError):                              # a device the job does not own: nothing is spawned
        sweep.run_sweep({"output_base_directory": str(tmp_path)}, conds[:2], [0, 2], run_fn=_report_visible_devices,
                        log=lambda *_: None)
    monkeypatch.delenv("CUDA_VISIBLE_DEVICES")
    assert sweep.pinned_device_env(3) == "3"


def _dp_worker(rank, world, port, out_dir):
    import torch.distributed as dist
    from hba import dp
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    groups = [("head", [torch.empty(10), torch.empty(3)]), ("block1", [torch.empty(7, 5)]), ("block0", [torch.empty(6)])]
    total, offsets, buckets = dp.layout_buckets(groups)
    flat = torch.zeros(total)
    red = dp.BucketAllReducer(dist)
    # parameters start equal on every rank after the broadcast
    w = torch.full((5,), float(rank + 1))
    red.broadcast_parameters([w])
    # local "gradients": rank r contributes (r + 1) * (bucket index + 1), scaled by 1/world up front
    for i, (name, s, e) in enumerate(buckets):
        flat[s:e] = dp.fold_world_size(torch.full((e - s,), float((rank + 1) * (i + 1))), world)
        red.on_bucket_ready(name, flat[s:e])
    launched = list(red.launched)
    red.wait()
    torch.save({"flat": flat, "w": w, "launched": launched, "buckets": buckets, "offsets": list(offsets.values())},
               os.path.join(out_dir, f"rank{rank}.pt"))
    dist.destroy_process_group()


def test_gradient_buckets_all_reduce_on_gloo_world_2(tmp_path):
    import torch.multiprocessing as mp
    port = 29650 + os.getpid() % 200
    mp.spawn(_dp_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r0, r1 = (torch.load(os.path.join(str(tmp_path), f"rank{r}.pt")) for r in (0, 1))
    assert torch.equal(r0["flat"], r1["flat"]) and torch.equal(r0["w"], r1["w"]) and float(r0["w"][0]) == 1.0
    assert r0["launched"] == ["head", "block1", "block0"]          # backward order = bucket order
    for i, (name, s, e) in enumerate(r0["buckets"]):
        assert s % 4 == 0 and torch.allclose(r0["flat"][s:e], torch.full((e - s,), 1.5 * (i + 1)))  # mean of ranks
    assert all(off % 4 == 0 for off, _ in r0["offsets"])            # every gradient 16-byte aligned


def test_bucket_reducer_can_be_switched_off_for_compute_only_timing():
    """bench.py times the multi-rank ViT step once more with the all-reduces left out (`enabled = False`): buckets
    are still announced in backward order, nothing is reduced, `wait()` has nothing to wait for."""
    from hba.dp import BucketAllReducer

    class _Dist:
        calls = 0

        def get_world_size(self, group=None):
            return 2

        def all_reduce(self, t, group=None, async_op=False):
            _Dist.calls += 1

            class _H:
                def wait(self_inner):
                    return None
            return _H()
    r = BucketAllReducer(_Dist())
    r.on_bucket_ready("head", torch.zeros(4))
    assert _Dist.calls == 1 and len(r.handles) == 1
    r.wait()
    r.enabled = False
    r.on_bucket_ready("block1", torch.zeros(4))
    assert _Dist.calls == 1 and r.handles == [] and r.launched == ["block1"]
    r.wait()
    assert r.launched == []


def _rsa_scale_worker(rank, world, port, root):
    """RSA at scale over DoRA checkpoints, sharded over 2 `gloo` ranks (hba.rsa_scale; the device pieces replaced by
    CPU stand-ins: embeddings derived from the loaded checkpoint, numpy / scipy for the RSA tail)."""
    import numpy as np
    import torch.distributed as dist
    from scipy.stats import spearmanr
    from hba import rsa_scale
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)

    class _Model(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.m = torch.nn.Parameter(torch.zeros(3))
    rng = np.random.default_rng(0)
    base = rng.standard_normal((12, 66))
    ref = rsa_scale.reference_rdm_from_targets(rng.standard_normal((12, 66)))
    iu = np.triu_indices(12, k=1)

    extra = rng.standard_normal((12, 66))

    def embed(model, loaders, device):
        with torch.no_grad():
            return torch.from_numpy(base * float(model.m[0]) + extra * 30.0 * float(model.m[1]))

    def evaluator(emb, want_rdm=False):
        rdm = 1 - np.corrcoef(emb.numpy())
        rho, p = spearmanr(ref[iu], rdm[iu])
        return float(rho), float(p), None
    files = rsa_scale.find_dora_checkpoints(root)
    out = rsa_scale.clip_rsa_over_checkpoints(_Model(), [], ref, files, "cpu", rank=rank, world_size=world,
                                              output_csv=os.path.join(root, "rsa.csv"), log=None, root=root,
                                              evaluator=evaluator, embed_fn=embed)
    torch.save({"out": out}, os.path.join(root, f"result{rank}.pt"))
    dist.destroy_process_group()


def test_rsa_at_scale_shards_checkpoints_over_gloo_world_2(tmp_path):
    import pandas as pd
    import torch.multiprocessing as mp
    root = str(tmp_path)
    want = {}
    for run, epochs in (("base/dora", (1, 2, 10)), ("out/random_target_e1_l2/dora_params_1", (3, 4))):
        os.makedirs(os.path.join(root, run))
        for e in epochs:
            torch.save({"m": torch.tensor([1.0 + 0.1 * e, 0.01 * e, 0.0])}, os.path.join(root, run, f"epoch{e}_dora_params.pth"))
            want[os.path.join(run, f"epoch{e}_dora_params.pth")] = e
    port = 29750 + os.getpid() % 200
    mp.spawn(_rsa_scale_worker, args=(2, port, root), nprocs=2, join=True)
    r0, r1 = (torch.load(os.path.join(root, f"result{r}.pt"), weights_only=False)["out"] for r in (0, 1))
    assert r1 is None                                        # only rank 0 holds the gathered rows
    rows, stats = r0
    assert [r["checkpoint"] for r in rows] == sorted(want)   # every file exactly once, whatever rank evaluated it
    assert [r["epoch"] for r in rows] == [want[r["checkpoint"]] for r in rows]
    assert sorted(s["checkpoints"] for s in stats) == [2, 3] and {s["rank"] for s in stats} == {0, 1}
    assert all(-1.0 <= r["behavioral_rsa_rho"] <= 1.0 for r in rows)
    assert len({round(r["behavioral_rsa_rho"], 12) for r in rows}) == len(rows)   # each from its own checkpoint
    df = pd.read_csv(os.path.join(root, "rsa.csv"))
    assert list(df.columns) == list(__import__("hba.rsa_scale", fromlist=["x"]).RESULT_COLUMNS) and len(df) == 5


def test_reference_arm_under_torchrun_two_ranks(tmp_path):
    """`bench.py --impl reference --gpus 2` launched the way the driver launches N > 1: rank 0 alone
    runs the CPU implementation and prints the JSON line, the other rank exits 0 without work."""
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="", OMP_NUM_THREADS="2")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", str(29850 + os.getpid() % 100), os.path.join(ROOT, "bench.py"), "--impl",
           "reference", "--gpus", "2", "--steps", "1", "--warmup", "0", "--backbone", "ViT-tiny/14",
           "--cpu-sample-images", "2"]
    p = subprocess.run(cmd, capture_output=True, text=True, env=env, timeout=600, cwd=ROOT)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [l for l in p.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["n_gpus"] == 2 and d["value"] > 0 and d["unit"] == "images/s"
    assert d["cpu_baseline"]["kind"] == "port" and d["e2e"]["h2d_bytes_per_step"] == 0


# ------------------------------------------------------------------------------- LEN's resume chain
def test_find_previous_run_dir_and_last_completed_epoch(tmp_path):
    """LEN:139-160 and LEN:188-221 (the reference keeps both inside `main()`; restated in hba.sweep)."""
    from hba import sweep
    base = str(tmp_path)
    for name in ("random_target_e1_l2", "random_target_e1_l10", "random_target_e1_l30", "random_target_e10_l5",
                 "label_shuffle_e1_l5", "random_target_e1_lx", "random_target_e11_l2"):
        os.makedirs(os.path.join(base, name))
    open(os.path.join(base, "random_target_e1_l20"), "w").close()          # a file, not a run directory
    find = sweep.find_previous_run_dir
    assert find(base, "random_target", 1, 20) == (os.path.join(base, "random_target_e1_l10"), 10)
    assert find(base, "random_target", 1, 50) == (os.path.join(base, "random_target_e1_l30"), 30)
    assert find(base, "random_target", 1, 2) == (None, None)                 # nothing shorter
    assert find(base, "random_target", 10, 50) == (os.path.join(base, "random_target_e10_l5"), 5)   # e1_ != e10_
    assert find(base, "label_shuffle", 1, 50) == (os.path.join(base, "label_shuffle_e1_l5"), 5)    # prefix must match
    assert find(base, "random_target", 2, 50) == (None, None)
    assert find(os.path.join(base, "missing"), "random_target", 1, 50) == (None, None)
    csv = os.path.join(base, "training_res.csv")
    assert sweep.last_completed_epoch(csv) == -1
    with open(csv, "w") as f:
        f.write("epoch,train_loss,test_loss,behavioral_rsa_rho,behavioral_rsa_p_value\n1,2.0,2.1,0.3,0.0\n"
                "2,1.9,2.0,0.31,0.0\n\nnot-a-number,1,1,1,1\n7,1.5,1.6,0.4,0.0\n")
    assert sweep.last_completed_epoch(csv) == 6                              # 1-indexed in the file


def test_apply_length_resume_branches(tmp_path):
    from hba import sweep
    base = {"output_base_directory": str(tmp_path), "perturb_type": "random_target"}
    logs = []
    # (3) nothing there: baseline epoch e-1
    cfg = sweep.condition_config(base, {"training_run": 8, "perturb_length": 5}, "length")
    assert sweep.apply_length_resume(cfg, logs.append) == "baseline" and cfg["resume_from_epoch"] == 7
    assert "resume_dora_parameters_path" not in cfg
    # a shorter neighbour directory WITHOUT the needed checkpoint (failed / still running): still baseline
    cfg10 = sweep.condition_config(base, {"training_run": 8, "perturb_length": 10}, "length")
    assert sweep.apply_length_resume(cfg10, logs.append) == "baseline" and cfg10["resume_from_epoch"] == 7
    assert "has no checkpoint of epoch 12" in logs[-1] and "shorter" in logs[-1]
    # (2) the shorter run finished its window: resume from its end, last_epoch = (e-1) + prev_length  (LEN:239-244)
    prev = os.path.join(str(tmp_path), "random_target_e8_l5")
    os.makedirs(os.path.join(prev, "dora_params_8"))
    open(os.path.join(prev, "dora_params_8", "epoch12_dora_params.pth"), "w").close()
    cfg10 = sweep.condition_config(base, {"training_run": 8, "perturb_length": 10}, "length")
    assert sweep.apply_length_resume(cfg10, logs.append) == "chain"
    assert cfg10["resume_from_epoch"] == 12 and cfg10["perturb_length"] == 10 and cfg10["training_run"] == 8
    assert cfg10["previous_training_res_path"] == os.path.join(prev, "training_res.csv")
    assert cfg10["resume_random_state_path"] == os.path.join(prev, "random_states_8")
    assert cfg10["resume_dora_parameters_path"] == os.path.join(prev, "dora_params_8")
    assert "resuming from epoch 13" in logs[-1]
    # (1) the run's own CSV exists: continue it from its own checkpoints  (LEN:229-236)
    with open(cfg10["training_res_path"], "w") as f:
        f.write("epoch,train_loss\n13,1.0\n14,0.9\n")
    cfg10 = sweep.condition_config(base, {"training_run": 8, "perturb_length": 10}, "length")
    assert sweep.apply_length_resume(cfg10) == "existing" and cfg10["resume_from_epoch"] == 14
    assert cfg10["previous_training_res_path"] == cfg10["training_res_path"]
    assert cfg10["resume_dora_parameters_path"] == cfg10["dora_parameters_path"]
    assert cfg10["resume_random_state_path"] == cfg10["random_state_path"]
    # first epoch: max(0, e-1)
    cfg = sweep.condition_config(base, {"training_run": 0, "perturb_length": 2}, "length")
    assert cfg["resume_from_epoch"] == 0


def test_chain_groups_and_costs():
    from hba import sweep
    grid = sweep.length_grid_conditions()
    groups = sweep.chain_groups(grid)
    assert len(groups) == 21 and sum(len(g) for g in groups) == 136
    for g in groups:
        assert len({c["training_run"] for c in g}) == 1
        lengths = [c["perturb_length"] for c in g]
        assert lengths == sorted(lengths)
    g1 = next(g for g in groups if g[0]["training_run"] == 1)
    indep = sum(sweep.expected_epochs(c) for c in g1)
    assert sweep.chain_cost(g1) == indep - (2 + 5 + 10 + 20 + 30 + 40)      # every window but the last is skipped once
    total_chain = sum(sweep.chain_cost(g) for g in groups)
    total_indep = sum(sweep.expected_epochs(c) for c in grid)
    assert total_chain < 0.86 * total_indep
    plan, loads = sweep.lpt_assign(groups, 8, cost=sweep.chain_cost)
    assert sum(len(p) for p in plan) == 21 and max(loads) < 1.1 * sum(loads) / 8


def _fake_length_condition(cfg):
    """Stand-in for run_behavioral_training on the 'length' layout: writes what a finished run leaves behind
    (result CSV rows and per-epoch DoRA checkpoints from the resume epoch to the end of the window + 2) and
    records how it was asked to resume."""
    e, length = cfg["training_run"], cfg["perturb_length"]
    if (e, length) == (3, 5):
        raise RuntimeError("condition (3, 5) fails")
    os.makedirs(cfg["dora_parameters_path"], exist_ok=True)
    first = cfg["resume_from_epoch"] + 1
    last = max(0, e - 1) + length + 2
    with open(cfg["training_res_path"], "a") as f:
        if f.tell() == 0:
            f.write("epoch,train_loss\n")
        for ep in range(first, last + 1):
            f.write(f"{ep},1.0\n")
            open(os.path.join(cfg["dora_parameters_path"], f"epoch{ep}_dora_params.pth"), "w").close()
    with open(os.path.join(cfg["output_dir"], "how.json"), "w") as f:
        json.dump({"pid": os.getpid(), "kind": cfg.get("hba_resume_kind"), "resume_from_epoch": cfg["resume_from_epoch"],
                   "resume_dora": cfg.get("resume_dora_parameters_path"), "t": __import__("time").time()}, f)


def test_chained_length_sweep_on_two_workers(tmp_path):
    from hba import sweep
    conds = [{"training_run": s, "perturb_length": l} for s in (1, 3, 7) for l in (2, 5, 10)]
    base = {"output_base_directory": str(tmp_path), "perturb_type": "random_target", "perturb_length": 1}
    logs = []
    res = sweep.run_sweep(base, conds, [None, None], layout="length", run_fn=_fake_length_condition, log=logs.append,
                          chain=True)
    assert [r["ok"] for r in res] == [not (c["training_run"] == 3 and c["perturb_length"] == 5) for c in conds]
    assert "8 successful, 1 failed" in logs[-1]
    how = {}
    for c in conds:
        p = os.path.join(str(tmp_path), f"random_target_e{c['training_run']}_l{c['perturb_length']}", "how.json")
        if os.path.exists(p):
            how[(c["training_run"], c["perturb_length"])] = json.load(open(p))
    for s in (1, 3, 7):
        chain = [how[(s, l)] for l in (2, 5, 10) if (s, l) in how]
        assert len({h["pid"] for h in chain}) == 1                            # one start epoch = one worker
        assert [h["t"] for h in chain] == sorted(h["t"] for h in chain)       # in increasing window length
    assert how[(1, 2)]["kind"] == "baseline" and how[(1, 2)]["resume_from_epoch"] == 0
    assert how[(1, 5)]["kind"] == "chain" and how[(1, 5)]["resume_from_epoch"] == 0 + 2
    assert how[(1, 10)]["kind"] == "chain" and how[(1, 10)]["resume_from_epoch"] == 0 + 5
    assert how[(7, 10)]["resume_from_epoch"] == 6 + 5
    assert how[(7, 10)]["resume_dora"].endswith(os.path.join("random_target_e7_l5", "dora_params_7"))
    # (3, 5) failed before writing anything: (3, 10) chains from the last FINISHED shorter window, (3, 2)
    assert how[(3, 10)]["kind"] == "chain" and how[(3, 10)]["resume_from_epoch"] == 2 + 2
    # independent mode keeps every condition on the baseline checkpoint
    res = sweep.run_sweep(dict(base, output_base_directory=os.path.join(str(tmp_path), "indep")), conds[:3], [None],
                          layout="length", run_fn=_fake_length_condition, log=logs.append)
    assert all(r["ok"] for r in res)
    h = json.load(open(os.path.join(str(tmp_path), "indep", "random_target_e1_l10", "how.json")))
    assert h["kind"] is None and h["resume_from_epoch"] == 0
    with pytest.raises(ValueError):
        sweep.run_sweep(base, conds, [None], layout="sweep", run_fn=_fake_length_condition, chain=True)


# ------------------------------------------------------------------------------- LEN executed
LEN_PATH = "/root/reference/Training/clip_behavioral_finetuning/length_experiments/clip_train_behavior_lengths.py"


def _run_len_main(argv, monkeypatch):
    """The reference's length-experiment driver, unmodified: `main()` with a command line, its
    `run_behavioral_training` replaced by a recorder -> the config dict it would train with."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("_ref_LEN", LEN_PATH)
    LEN = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(LEN)          # resolves `functions.cvpr_train_behavior_things_pipeline` to this repo's alias
    seen = []
    monkeypatch.setattr(LEN, "run_behavioral_training", lambda cfg: seen.append(dict(cfg)))
    monkeypatch.setattr(sys, "argv", ["clip_train_behavior_lengths.py"] + argv)
    LEN.main()
    return seen[0]


@pytest.mark.skipif(not os.path.exists(LEN_PATH), reason="reference not mounted")
def test_length_resume_decisions_equal_the_reference_driver_executed(tmp_path, monkeypatch):
    """hba.sweep.condition_config('length') + apply_length_resume against LEN's own `main()` (LEN:85-253) on the
    same directory states: fresh start, shorter finished neighbour (chain), own CSV present (continue)."""
    from hba import sweep
    import functions.cvpr_train_behavior_things_pipeline as alias
    import functions.new_cvpr_train_behavior_things_pipeline as NEW
    assert alias.run_behavioral_training is NEW.run_behavioral_training          # LEN:1 resolves here
    out = str(tmp_path / "out")
    os.makedirs(out)
    keys = ("training_run", "perturb_length", "perturb_type", "resume_from_epoch", "output_dir", "checkpoint_path",
            "training_res_path", "dora_parameters_path", "random_state_path", "previous_training_res_path",
            "resume_random_state_path", "resume_dora_parameters_path")

    def both(e, length):
        argv = ["--perturb_epoch", str(e), "--perturb_length", str(length), "--output_dir", f"random_target_e{e}_l{length}",
                "--baseline_dora_directory", "b/dora", "--baseline_random_state_path", "b/rand",
                "--baseline_split_indices_path", "b/split.pth", "--output_base_directory", out]
        ref_cfg = _run_len_main(argv, monkeypatch)
        cfg = sweep.condition_config({"output_base_directory": out, "perturb_type": "random_target"},
                                     {"training_run": e, "perturb_length": length}, "length")
        kind = sweep.apply_length_resume(cfg)
        return {k: ref_cfg.get(k) for k in keys}, {k: cfg.get(k) for k in keys}, kind

    # (3) nothing on disk: baseline epoch e-1
    ref_cfg, cfg, kind = both(8, 5)
    assert kind == "baseline" and cfg == ref_cfg and cfg["resume_from_epoch"] == 7
    # (2) the window-5 run has finished its window (checkpoint of epoch 7 + 5 present): the window-10 run chains
    os.makedirs(os.path.join(out, "random_target_e8_l5", "dora_params_8"), exist_ok=True)
    open(os.path.join(out, "random_target_e8_l5", "dora_params_8", "epoch12_dora_params.pth"), "w").close()
    os.makedirs(os.path.join(out, "random_target_e80_l2"))                       # e8_ must not match e80_
    os.makedirs(os.path.join(out, "label_shuffle_e8_l7"))                        # other perturbation type
    ref_cfg, cfg, kind = both(8, 10)
    assert kind == "chain" and cfg == ref_cfg and cfg["resume_from_epoch"] == 12
    # (1) the run's own CSV exists: continue after its last completed epoch
    with open(os.path.join(out, "random_target_e8_l10", "training_res.csv"), "w") as f:
        f.write("epoch,train_loss\n13,1.0\n14,0.9\n")
    ref_cfg, cfg, kind = both(8, 10)
    assert kind == "existing" and cfg == ref_cfg and cfg["resume_from_epoch"] == 14
    # deliberate difference: a shorter neighbour WITHOUT its checkpoint (failed / still running) - the reference
    # chains to it regardless (and would then train from un-restored state); here it is stepped over
    os.makedirs(os.path.join(out, "random_target_e30_l2"))
    ref_cfg, cfg, kind = both(30, 5)
    assert ref_cfg["resume_from_epoch"] == 31 and kind == "baseline" and cfg["resume_from_epoch"] == 29


@pytest.mark.skipif(not os.path.exists(LEN_PATH), reason="reference not mounted")
def test_length_resume_decisions_on_random_directory_states(tmp_path, monkeypatch):
    """Seeded random trees of finished runs (every run holds all its checkpoints, so the one deliberate
    difference - stepping over a neighbour without its checkpoint - cannot trigger): the config built here equals
    the one LEN's own `main()` builds, for every query."""
    import random as pyrandom
    from hba import sweep
    rng = pyrandom.Random(7)
    keys = ("training_run", "perturb_length", "resume_from_epoch", "training_res_path", "dora_parameters_path",
            "random_state_path", "previous_training_res_path", "resume_random_state_path", "resume_dora_parameters_path")
    kinds = []
    for trial in range(4):
        out = str(tmp_path / f"tree{trial}")
        os.makedirs(out)
        for _ in range(rng.randint(3, 9)):
            t, e, l = rng.choice(["random_target", "label_shuffle"]), rng.choice([1, 2, 10, 11, 12, 21]), rng.choice([2, 5, 10, 20])
            d = os.path.join(out, f"{t}_e{e}_l{l}", f"dora_params_{e}")
            os.makedirs(d, exist_ok=True)
            for ep in range(max(0, e - 1) + 1, max(0, e - 1) + l + 3):
                open(os.path.join(d, f"epoch{ep}_dora_params.pth"), "w").close()
            if rng.random() < 0.3:      # some runs are also resumable in place
                with open(os.path.join(out, f"{t}_e{e}_l{l}", "training_res.csv"), "w") as f:
                    f.write("epoch,train_loss\n" + "".join(f"{ep},1.0\n" for ep in range(1, max(0, e - 1) + rng.randint(1, l) + 1)))
        for _ in range(6):
            t, e, l = rng.choice(["random_target", "label_shuffle"]), rng.choice([1, 2, 10, 11, 12, 21]), rng.choice([2, 5, 10, 20, 50])
            argv = ["--perturb_type", t, "--perturb_epoch", str(e), "--perturb_length", str(l), "--output_dir", f"{t}_e{e}_l{l}",
                    "--baseline_dora_directory", "b/dora", "--baseline_random_state_path", "b/rand",
                    "--baseline_split_indices_path", "b/split.pth", "--output_base_directory", out]
            existed = os.path.isdir(os.path.join(out, f"{t}_e{e}_l{l}"))
            ref_cfg = _run_len_main(argv, monkeypatch)
            cfg = sweep.condition_config({"output_base_directory": out, "perturb_type": t},
                                         {"training_run": e, "perturb_length": l}, "length")
            kinds.append(sweep.apply_length_resume(cfg))
            assert {k: cfg.get(k) for k in keys} == {k: ref_cfg.get(k) for k in keys}, (trial, t, e, l)
            if not existed:      # both sides create the queried run's directory: keep the tree to finished runs only
                __import__("shutil").rmtree(os.path.join(out, f"{t}_e{e}_l{l}"))
    assert {"baseline", "chain", "existing"} <= set(kinds), kinds      # the random states reached every branch


# ------------------------------------------------------------------------------- SWEEP executed
SWEEP_PATH = "/root/reference/Training/clip_behavioral_finetuning/uniform_sweep/clip_train_behavior_sweep.py"


@pytest.mark.skipif(not os.path.exists(SWEEP_PATH), reason="reference not mounted")
def test_single_epoch_layout_equals_the_reference_sweep_driver_executed(monkeypatch):
    """SWEEP's own `main()` (SWEEP:111-237) with a recording `run_behavioral_training`: per-condition paths, resume
    epoch and loop behaviour (a failing run is counted and the loop goes on) against `condition_config('sweep')` /
    `run_sweep`.  The driver hard-codes an output tree under /home: directory creation and its log file are
    intercepted, nothing is written outside the test."""
    import importlib.util
    import logging
    from hba import sweep
    spec = importlib.util.spec_from_file_location("_ref_SWEEP", SWEEP_PATH)
    SWEEP = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(SWEEP)
    seen, made = [], []

    def record(cfg):
        seen.append(dict(cfg))
        if cfg["training_run"] == 25:
            raise RuntimeError("run 25 fails")                    # SWEEP:215-223: logged, counted, loop continues

    monkeypatch.setattr(SWEEP, "run_behavioral_training", record)
    monkeypatch.setattr(os, "makedirs", lambda p, *a, **k: made.append(p))
    monkeypatch.setattr(logging, "FileHandler", lambda *a, **k: logging.NullHandler())
    SWEEP.main()
    assert [c["training_run"] for c in seen] == [15, 25, 35, 70]   # the shipped training_order; all four attempted
    keys = ("training_run", "resume_from_epoch", "checkpoint_path", "training_res_path", "dora_parameters_path",
            "random_state_path", "perturb_length", "perturb_type", "perturb_seed", "output_base_directory")
    base = {k: v for k, v in seen[0].items() if k not in ("training_run", "resume_from_epoch", "checkpoint_path",
                                                           "training_res_path", "dora_parameters_path", "random_state_path")}
    for ref_cfg in seen:
        cfg = sweep.condition_config(base, {"training_run": ref_cfg["training_run"], "perturb_length": 1}, "sweep")
        assert {k: cfg[k] for k in keys} == {k: ref_cfg[k] for k in keys}
        assert os.path.join(base["output_base_directory"], f"training_run{ref_cfg['training_run']}") in made
    assert all(p.startswith(base["output_base_directory"]) for p in made)
    # the midpoint order helper the driver ships covers every epoch once (SWEEP:8-60); run_sweep orders by cost instead
    order = SWEEP.generate_midpoint_order(1, 98)
    assert sorted(order) == list(range(1, 99)) == [c["training_run"] for c in sweep.single_epoch_conditions(1, 98)]


# ------------------------------------------------------------------------------- config contract of the drivers
BDRV_PATH = "/root/reference/Training/clip_behavioral_finetuning/baseline/clip_train_behavior_baseline.py"


def _required_config_keys(*modules):
    """Keys read as `config['k']` (no default) anywhere in the given modules' source."""
    import ast
    import inspect
    keys, guarded = set(), set()
    for mod in modules:
        for node in ast.walk(ast.parse(inspect.getsource(mod))):
            if (isinstance(node, ast.Subscript) and isinstance(node.value, ast.Name) and node.value.id in ("config", "cfg")
                    and isinstance(node.slice, ast.Constant) and isinstance(node.slice.value, str)
                    and isinstance(node.ctx, ast.Load)):
                keys.add(node.slice.value)
            if (isinstance(node, ast.Call) and isinstance(node.func, ast.Attribute) and node.func.attr == "get"
                    and isinstance(node.func.value, ast.Name) and node.func.value.id in ("config", "cfg")
                    and node.args and isinstance(node.args[0], ast.Constant)):
                guarded.add(node.args[0].value)      # `if config.get('k'): ... config['k']` is an optional key
    return keys - guarded


@pytest.mark.skipif(not os.path.exists(BDRV_PATH), reason="reference not mounted")
def test_pipelines_require_only_keys_the_reference_drivers_provide(tmp_path, monkeypatch):
    """Drop-in contract of the config dict (SURVEY 5 'Config / flags'): every key the drop-in pipelines read
    without a default is in the dict the reference's own drivers build (BDRV / SWEEP / LEN `main()` executed with a
    recording `run_behavioral_training`); keys this repo adds (`hba_*`) are optional."""
    import importlib.util
    import logging
    import functions._pipeline_core as core
    import functions.cvpr_train_behavior_things_pipeline_baseline as BASE
    import functions.new_cvpr_train_behavior_things_pipeline as NEW

    def driver_config(path, argv=None):
        spec = importlib.util.spec_from_file_location("_ref_driver", path)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        seen = []
        monkeypatch.setattr(mod, "run_behavioral_training", lambda cfg: seen.append(dict(cfg)))
        monkeypatch.setattr(sys, "argv", [os.path.basename(path)] + (argv or []))
        mod.main()
        return seen[0]

    bdrv = driver_config(BDRV_PATH)
    len_cfg = driver_config(LEN_PATH, ["--perturb_epoch", "3", "--perturb_length", "2", "--output_dir", "random_target_e3_l2",
                                       "--baseline_dora_directory", "b", "--baseline_random_state_path", "r",
                                       "--baseline_split_indices_path", "s", "--output_base_directory", str(tmp_path)])
    with monkeypatch.context() as m:
        m.setattr(os, "makedirs", lambda p, *a, **k: None)
        m.setattr(logging, "FileHandler", lambda *a, **k: logging.NullHandler())
        sweep_cfg = driver_config(SWEEP_PATH)
    base_needs = _required_config_keys(BASE)
    new_needs = _required_config_keys(NEW)
    shared_needs = _required_config_keys(core)
    assert base_needs and new_needs
    # (shared helpers also serve the perturbation pipeline: keys only its drivers provide are not needed by BDRV runs)
    base_missing = (base_needs | shared_needs) - set(bdrv)
    assert base_missing <= {k for k in shared_needs if k in sweep_cfg}, sorted(base_missing)
    for name, cfg in (("SWEEP", sweep_cfg), ("LEN", len_cfg)):
        missing = (new_needs | shared_needs) - set(cfg)
        assert not missing, (name, sorted(missing))
    assert not any(k.startswith("hba_") for k in base_needs | new_needs | shared_needs)      # our additions are .get() only


# ------------------------------------------------------------------------------- bench.py: scheduler-run slice at N > 1
_FAKE_GRID_TOOL = '''
import argparse, json, os, subprocess, sys, time
ap = argparse.ArgumentParser()
for flag in ("--kind", "--gpus", "--per-gpu", "--max-start", "--batch-size", "--backbone", "--root", "--out"):
    ap.add_argument(flag)
a = ap.parse_args()
assert "RANK" not in os.environ and "WORLD_SIZE" not in os.environ   # the tool must not look like a rank of the job
if os.environ.get("FAKE_TOOL_MODE") == "hang":
    child = subprocess.Popen([sys.executable, "-c", "import time; time.sleep(600)"])   # a sweep worker of its own
    with open(os.environ["FAKE_TOOL_PIDS"], "w") as f:
        f.write(f"{os.getpid()} {child.pid}")
    time.sleep(600)
time.sleep(1.5)
gpus = a.gpus.split(",")
json.dump({"conditions": len(gpus) * int(a.per_gpu), "conditions_per_hour": 123.0, "gpus": gpus, "wall_s": 1.5,
           "per_condition": [{"epochs_trained": 7}] * (len(gpus) * int(a.per_gpu))}, open(a.out, "w"))
'''


def _fake_bench_root(tmp_path):
    root = os.path.join(str(tmp_path), "fake_root")
    os.makedirs(os.path.join(root, "tools"))
    with open(os.path.join(root, "tools", "grid_sweep_bench.py"), "w") as f:
        f.write(_FAKE_GRID_TOOL)
    return root


_SCHED_RANK_SCRIPT = '''
import argparse, os, sys, time
import torch
import torch.distributed as dist
sys.path.insert(0, {root!r})
import bench
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo")                      # env:// rendezvous through torchrun's agent store, as the driver's launch
bench.ROOT = {fake_root!r}
args = argparse.Namespace(sweep_per_gpu=3, batch=32, backbone="ViT-tiny/14", sweep_timeout=120.0)
t0 = time.time()
res = bench.measure_sweep_scheduler(args, world, rank, dist)
torch.save({{"res": res, "t0": t0, "t1": time.time()}}, os.path.join({out_dir!r}, f"sched{{rank}}.pt"))
dist.barrier()                 # the group is still usable by every rank afterwards (the ViT section follows)
dist.destroy_process_group()
'''


def test_bench_scheduler_slice_runs_on_rank0_while_other_ranks_wait_on_the_store(tmp_path):
    """bench.py at N > 1 (the driver's 2 / 4 / 8-GPU runs, launched through torch.distributed.run): rank 0 alone
    launches the sweep tool over all N GPUs, the other ranks block on the rendezvous store (no collective in flight)
    and return only once the slice is done."""
    script = os.path.join(str(tmp_path), "sched_rank.py")
    with open(script, "w") as f:
        f.write(_SCHED_RANK_SCRIPT.format(root=ROOT, fake_root=_fake_bench_root(tmp_path), out_dir=str(tmp_path)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", str(29450 + os.getpid() % 200), script]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT,
                       env=dict(os.environ, CUDA_VISIBLE_DEVICES=""))
    assert p.returncode == 0, p.stderr[-3000:]
    r0, r1 = (torch.load(os.path.join(str(tmp_path), f"sched{r}.pt"), weights_only=False) for r in (0, 1))
    assert r1["res"] is None
    res = r0["res"]
    assert "error" not in res and res["conditions"] == 6 and res["gpus"] == ["0", "1"]
    assert res["unit"] == "conditions/h" and res["epochs_per_condition"] == [7] * 6 and "per_condition" not in res
    assert res["tool_wall_s"] >= 1.5
    assert r1["t1"] >= r0["t0"] + 1.5      # rank 1 was released by rank 0's store key, not before the tool ended


def test_bench_scheduler_slice_timeout_kills_the_tool_and_its_workers(tmp_path, monkeypatch):
    """A slice that overruns --sweep-timeout is killed as a process group (tool + the workers it spawned) and
    reported as an error; the bench line goes on."""
    import argparse
    import time
    monkeypatch.syspath_prepend(ROOT)
    import bench
    monkeypatch.setattr(bench, "ROOT", _fake_bench_root(tmp_path))
    pids_file = os.path.join(str(tmp_path), "pids")
    monkeypatch.setenv("FAKE_TOOL_MODE", "hang")
    monkeypatch.setenv("FAKE_TOOL_PIDS", pids_file)
    for k in ("RANK", "WORLD_SIZE"):
        monkeypatch.delenv(k, raising=False)
    args = argparse.Namespace(sweep_per_gpu=1, batch=32, backbone="ViT-tiny/14", sweep_timeout=3.0)
    t0 = time.time()
    res = bench.measure_sweep_scheduler(args, 1, 0, None)
    assert time.time() - t0 < 60
    assert "timed out after 3.0 s" in res["error"]
    # --time-budget: the slice gets what is left of the run's budget minus the reserve for the sections after it
    monkeypatch.setattr(bench, "SWEEP_MIN_S", 1.0)
    args.sweep_timeout = 300.0
    args.time_budget = (time.perf_counter() - bench.T_START) + bench.SWEEP_RESERVE_S + 3.0
    t0 = time.time()
    res = bench.measure_sweep_scheduler(args, 1, 0, None)
    limit = float(res["error"].split("timed out after ")[1].split(" s")[0])
    assert 1.0 <= limit <= 3.0 and 1.0 <= time.time() - t0 < 30     # what was left of the budget, not --sweep-timeout
    args.time_budget = 1.0                                  # budget already spent: the floor applies
    t0 = time.time()
    res = bench.measure_sweep_scheduler(args, 1, 0, None)
    assert time.time() - t0 < 30 and "timed out after 1.0 s" in res["error"]
    tool_pid, child_pid = (int(x) for x in open(pids_file).read().split())
    time.sleep(0.5)
    for pid in (tool_pid, child_pid):
        try:                       # (a zombie that init has not reaped yet still answers signal 0: read its state)
            state = open(f"/proc/{pid}/stat").read().rsplit(")", 1)[1].split()[0]
        except FileNotFoundError:
            state = "gone"
        assert state in ("gone", "Z"), f"pid {pid} survived the timeout in state {state}"


# ------------------------------------------------------------------------------- ViT data parallel, the real trainer
def _vit_dp_worker(rank, world, port, out_dir):
    """hba.vit.DataParallelTrainer itself (engine forward / backward, bucket all-reduces announced by the backward
    pass, fused SGD) on a `gloo` group, libhba served by its CPU restatement."""
    import torch.distributed as dist
    import hba
    from hba import vit
    from oracle.libhba_ref import emulated_device
    torch.set_num_threads(2)
    if world > 1:
        dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    hba.set_precision("fp32")
    torch.manual_seed(100 + rank)                      # ranks start from DIFFERENT weights: the broadcast must fix that
    model = vit.create_model("vit_tiny_test", num_classes=10)
    g = torch.Generator().manual_seed(0)
    images = torch.randn(8, 3, 224, 224, generator=g)
    labels = torch.randint(0, 10, (8,), generator=g)
    share = slice(rank * 8 // world, (rank + 1) * 8 // world)
    with emulated_device():
        tr = vit.DataParallelTrainer(model, lr=0.05, momentum=0.9, weight_decay=1e-4, use_graph=False)
        if world > 1:
            tr.broadcast_parameters()
        else:
            torch.manual_seed(100)                    # (the single-rank arm: rank 0's initial weights)
            ref0 = vit.create_model("vit_tiny_test", num_classes=10)
            model.load_state_dict(ref0.state_dict())
        losses = []
        for _ in range(3):
            loss, hits = tr.step(images[share], labels[share])
            losses.append(float(loss))
        launched = list(tr.reducer.launched)
    torch.save({"params": {k: v.detach().clone() for k, v in model.state_dict().items()}, "losses": losses,
                "world": tr.world}, os.path.join(out_dir, f"vit_w{world}_r{rank}.pt"))
    if world > 1:
        dist.destroy_process_group()


def test_vit_data_parallel_trainer_world_2_equals_one_rank_on_the_full_batch(tmp_path):
    """SURVEY 8e, ViT-B/16 row (VIT:287 DistributedDataParallel): two ranks with half the batch each end three SGD steps
    with the parameters one rank reaches on the whole batch (mean over the GLOBAL batch: 1/world folded into
    dL/dlogits, SUM all-reduce per block bucket), after a broadcast made their different initial weights equal."""
    import torch.multiprocessing as mp
    port = 29350 + os.getpid() % 200
    mp.spawn(_vit_dp_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    _vit_dp_worker(0, 1, 0, str(tmp_path))
    r0, r1, one = (torch.load(os.path.join(str(tmp_path), f)) for f in ("vit_w2_r0.pt", "vit_w2_r1.pt", "vit_w1_r0.pt"))
    assert r0["world"] == 2 and one["world"] == 1
    worst = 0.0
    for k, v in one["params"].items():
        assert torch.equal(r0["params"][k], r1["params"][k]), k            # the ranks stay bit-identical to each other
        scale = float(v.abs().max().clamp_min(1e-6))
        worst = max(worst, float((r0["params"][k] - v).abs().max()) / scale)
    assert worst < 2e-4, worst                                             # = one rank on the full batch (fp32 re-association)
    # the global loss is the mean of the two half-batch losses
    for a, b, c in zip(r0["losses"], r1["losses"], one["losses"]):
        assert abs(0.5 * (a + b) - c) < 2e-4 * max(1.0, abs(c))
