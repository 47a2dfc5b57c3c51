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
