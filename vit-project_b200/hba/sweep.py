"""Multi-GPU perturbation sweeps: independent conditions, one (or more) per GPU, no collective.

Host-side mirror of the reference's two sweep drivers (pure Python, so is this):
  SWEEP = Training/clip_behavioral_finetuning/uniform_sweep/clip_train_behavior_sweep.py
          (single-epoch perturbations: `training_run` e in 1..98, resume from baseline epoch e-1,
          per-run directory `training_run{e}` and file names of SWEEP:198-207, failures logged and the
          loop continues, SWEEP:209-223)
  LEN   = Training/clip_behavioral_finetuning/length_experiments/clip_train_behavior_lengths.py
          (one process per (perturb_epoch, perturb_length) condition, paths of LEN:128-137)
and of the 136-condition grid shipped under Data/clip_results/perturb_length_experiments_* (SURVEY 8d).

The reference walks its conditions sequentially on one GPU.  Conditions share nothing but the read-only
baseline checkpoints, so here a pool of worker processes - one per GPU, pinned through
CUDA_VISIBLE_DEVICES before CUDA is initialised - pulls conditions from a queue, longest expected run
first (LPT), and calls the unmodified `run_behavioral_training(config)` with `config['cuda'] = 0`.
Inside a worker the frozen CLIP, its staged weights and the resident image store survive from condition
to condition (functions._pipeline_core.load_clip_to_cpu / _STORES).
"""
from __future__ import annotations

import copy
import heapq
import os
import time
import traceback

GRID_STARTS_FULL = (1, 2, 3, 6, 7, 8, 10, 20, 30, 40, 50, 60, 70, 80, 90)
GRID_LENGTHS_FULL = (2, 5, 10, 20, 30, 40, 50)
GRID_STARTS_PARTIAL = (13, 16, 19, 58, 94)
GRID_LENGTHS_PARTIAL = (5, 10, 20, 30, 40, 50)


def single_epoch_conditions(start=1, end=98):
    """SWEEP:192-207: one condition per perturbed epoch, window length 1."""
    return [{"training_run": e, "perturb_length": 1} for e in range(start, end + 1)]


def length_grid_conditions():
    """The 136 (start epoch, window length) pairs of the shipped length experiments (SURVEY 8d
    config 4): 15 starts x 7 lengths + 5 starts x 6 lengths + (22, 5)."""
    pairs = [(s, l) for s in GRID_STARTS_FULL for l in GRID_LENGTHS_FULL]
    pairs += [(s, l) for s in GRID_STARTS_PARTIAL for l in GRID_LENGTHS_PARTIAL]
    pairs.append((22, 5))
    return [{"training_run": s, "perturb_length": l} for s, l in sorted(pairs)]


def expected_epochs(cond, horizon=110):
    """Cost model for scheduling: epochs a condition trains before early stopping.  A run resumes at
    epoch training_run-1, the patience counter is frozen inside the window (NEW:1049-1063) and the
    baseline stops at ~epoch 90-110 (MAINLOG), so later starts are shorter and longer windows longer."""
    resume = max(0, cond["training_run"] - 1)
    return max(1, horizon - resume) + cond.get("perturb_length", 1)


def lpt_order(conditions, cost=expected_epochs):
    """Longest-processing-time-first order (ties keep the input order): with a shared queue this is
    the classic 4/3-approximation of the makespan."""
    return [c for _, _, c in sorted(((-cost(c), i, c) for i, c in enumerate(conditions)))]


def lpt_assign(conditions, n_workers, cost=expected_epochs):
    """Static LPT assignment -> (per-worker lists, per-worker total cost); what the dynamic queue does
    when the cost model is exact.  Used for planning / tests."""
    heap = [(0.0, w) for w in range(n_workers)]
    heapq.heapify(heap)
    plan = [[] for _ in range(n_workers)]
    for c in lpt_order(conditions, cost):
        load, w = heapq.heappop(heap)
        plan[w].append(c)
        heapq.heappush(heap, (load + cost(c), w))
    return plan, [sum(cost(c) for c in p) for p in plan]


def condition_config(base_config, cond, layout="sweep"):
    """Per-condition config dict.  layout 'sweep' = SWEEP:198-207 (training_run{e}/...), 'length' =
    LEN:128-137 ({perturb_type}_e{e}_l{len}/training_res.csv ...).  Every condition resumes from the
    BASELINE checkpoint of epoch training_run-1, which makes the conditions independent."""
    cfg = copy.copy(base_config)
    e, length = int(cond["training_run"]), int(cond.get("perturb_length", base_config.get("perturb_length", 1)))
    cfg["training_run"], cfg["perturb_length"] = e, length
    cfg["resume_from_epoch"] = max(0, e - 1)
    base = base_config["output_base_directory"]
    if layout == "sweep":
        d = os.path.join(base, f"training_run{e}")
        cfg["checkpoint_path"] = os.path.join(d, f"model_checkpoint_run{e}.pth")
        cfg["training_res_path"] = os.path.join(d, f"training_res_run{e}.csv")
        cfg["dora_parameters_path"] = os.path.join(d, f"dora_params_run{e}")
        cfg["random_state_path"] = os.path.join(d, f"random_states_run{e}")
    elif layout == "length":
        d = os.path.join(base, f"{base_config.get('perturb_type', 'random_target')}_e{e}_l{length}")
        cfg["output_dir"] = d
        cfg["checkpoint_path"] = os.path.join(d, f"model_checkpoint_{e}.pth")
        cfg["training_res_path"] = os.path.join(d, "training_res.csv")
        cfg["dora_parameters_path"] = os.path.join(d, f"dora_params_{e}")
        cfg["random_state_path"] = os.path.join(d, f"random_states_{e}")
    else:
        raise ValueError(f"unknown layout {layout!r}")
    os.makedirs(d, exist_ok=True)
    return cfg


def _default_run_fn(config):
    from functions.new_cvpr_train_behavior_things_pipeline import run_behavioral_training
    return run_behavioral_training(config)


def _worker(worker_id, device_id, tasks, results, base_config, layout, run_fn):
    # pin the GPU before anything initialises CUDA in this process
    if device_id is not None:
        os.environ["CUDA_VISIBLE_DEVICES"] = str(device_id)
    run_fn = run_fn or _default_run_fn
    while True:
        item = tasks.get()
        if item is None:
            break
        idx, cond = item
        t0 = time.time()
        try:
            cfg = condition_config(base_config, cond, layout)
            if device_id is not None:
                cfg["cuda"] = 0   # the only visible device (NEW:1137-1144)
            run_fn(cfg)
            results.put((idx, worker_id, True, time.time() - t0, ""))
        except Exception as exc:  # SWEEP:215-223: log, count, continue with the next condition
            results.put((idx, worker_id, False, time.time() - t0, f"{exc}\n{traceback.format_exc()}"))


def run_sweep(base_config, conditions, devices, layout="sweep", run_fn=None, cost=expected_epochs, log=print):
    """Runs `conditions` on one worker process per entry of `devices` (CUDA device indices; None
    entries run without pinning, for CPU tests).  Returns a list of result dicts in condition order:
    {condition, worker, ok, seconds, error}."""
    import multiprocessing as mp
    ctx = mp.get_context("spawn")
    tasks, results = ctx.Queue(), ctx.Queue()
    order = lpt_order(list(conditions), cost)
    index_of = {id(c): i for i, c in enumerate(conditions)}
    for c in order:
        tasks.put((index_of[id(c)], c))
    for _ in devices:
        tasks.put(None)
    procs = [ctx.Process(target=_worker, args=(w, dev, tasks, results, base_config, layout, run_fn), daemon=False)
             for w, dev in enumerate(devices)]
    t0 = time.time()
    for p in procs:
        p.start()
    out = [None] * len(conditions)
    done = 0
    while done < len(conditions):
        try:
            idx, worker, ok, secs, err = results.get(timeout=5)
        except Exception:
            if not any(p.is_alive() for p in procs):
                break
            continue
        out[idx] = {"condition": conditions[idx], "worker": worker, "ok": ok, "seconds": secs, "error": err}
        done += 1
        c = conditions[idx]
        log(f"[{done}/{len(conditions)}] run {c['training_run']} len {c.get('perturb_length', 1)} on worker "
            f"{worker}: {'ok' if ok else 'FAILED'} in {secs:.1f}s" + ("" if ok else f" - {err.splitlines()[0]}"))
    for p in procs:
        p.join()
    for i, r in enumerate(out):
        if r is None:   # a worker died before reporting (e.g. killed): count as failed, as SWEEP would
            out[i] = {"condition": conditions[i], "worker": None, "ok": False, "seconds": 0.0,
                      "error": "worker process exited without reporting"}
    wall = time.time() - t0
    n_ok = sum(r["ok"] for r in out)
    log(f"sweep finished: {n_ok} successful, {len(out) - n_ok} failed, {wall:.1f}s wall, "
        f"{3600.0 * n_ok / max(wall, 1e-9):.1f} conditions/hour on {len(devices)} worker(s)")
    return out
