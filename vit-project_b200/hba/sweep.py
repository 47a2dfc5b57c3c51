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


def last_completed_epoch(training_res_path):
    """LEN:139-160: the largest (1-indexed) epoch value in an existing result CSV, as a 0-indexed epoch;
    -1 when there is no usable row."""
    import csv
    last = -1
    if not os.path.exists(training_res_path):
        return last
    try:
        with open(training_res_path, "r") as f:
            reader = csv.reader(f)
            next(reader, None)
            for row in reader:
                if row:
                    try:
                        last = max(last, int(row[0]) - 1)
                    except (ValueError, IndexError):
                        continue
    except OSError:
        return -1
    return last


def find_previous_run_dir(base_dir, perturb_type, start_epoch, current_length):
    """LEN:188-221: among the run directories under `base_dir` with the same start epoch token `e{start}_`
    (and the same perturbation prefix), the one with the largest window length `_l{n}` below
    `current_length` -> (path, length) or (None, None)."""
    if not os.path.isdir(base_dir):
        return None, None
    candidates = []
    for name in os.listdir(base_dir):
        full = os.path.join(base_dir, name)
        if not os.path.isdir(full) or f"e{start_epoch}_" not in name:
            continue
        if perturb_type in ("random_target", "label_shuffle") and not name.startswith(perturb_type):
            continue
        length = next((int(p[1:]) for p in name.split("_") if p.startswith("l") and p[1:].isdigit()), None)
        if length is not None and length < current_length:
            candidates.append((length, full))
    if not candidates:
        return None, None
    length, path = max(candidates, key=lambda t: t[0])
    return path, length


def apply_length_resume(cfg, log=None):
    """The resume decisions of LEN:222-253 on a 'length'-layout config, in the reference's order:
    (1) the run's own result CSV exists -> continue after its last completed epoch from its own
        checkpoints; (2) a finished run with the same start epoch and a shorter window exists -> resume from
        the end of THAT window (both runs are identical up to there), `last_epoch = max(0, e-1) + prev_length`;
    (3) otherwise from the baseline checkpoint of epoch e-1 (what `condition_config` set).
    Beyond the reference, (2) only accepts a shorter run that really holds the DoRA checkpoint of `last_epoch`
    (a failed or still-running neighbour is stepped over: next shorter window, finally (3)).  Returns 'existing' | 'chain' |
    'baseline'."""
    say = log or (lambda *_: None)
    e, length = int(cfg["training_run"]), int(cfg["perturb_length"])
    done = last_completed_epoch(cfg["training_res_path"])
    if done >= 0:
        cfg["resume_from_epoch"] = done + 1
        cfg["previous_training_res_path"] = cfg["training_res_path"]
        cfg["resume_random_state_path"] = cfg["random_state_path"]
        cfg["resume_dora_parameters_path"] = cfg["dora_parameters_path"]
        say(f"Detected existing training run. Resuming from epoch {done + 2}")
        return "existing"
    below = length
    while True:   # the reference takes the longest shorter window; step down past neighbours without a checkpoint
        prev_dir, prev_length = find_previous_run_dir(cfg["output_base_directory"], cfg.get("perturb_type"), e, below)
        if prev_dir is None:
            break
        last_epoch = max(0, e - 1) + prev_length
        dora = os.path.join(prev_dir, f"dora_params_{e}", f"epoch{last_epoch}_dora_params.pth")
        if os.path.exists(dora):
            cfg["resume_from_epoch"] = last_epoch
            cfg["previous_training_res_path"] = os.path.join(prev_dir, "training_res.csv")
            cfg["resume_random_state_path"] = os.path.join(prev_dir, f"random_states_{e}")
            cfg["resume_dora_parameters_path"] = os.path.join(prev_dir, f"dora_params_{e}")
            say(f"Detected previous run at '{prev_dir}' with length {prev_length}; resuming from epoch {last_epoch + 1}")
            return "chain"
        say(f"Previous run at '{prev_dir}' has no checkpoint of epoch {last_epoch}; looking for a shorter one")
        below = prev_length
    return "baseline"


def chain_groups(conditions):
    """Conditions of the length grid grouped by start epoch, each group in increasing window length: the
    order in which LEN's shorter -> longer resume chain (LEN:188-253) can be used.  A group runs on ONE worker."""
    groups = {}
    for c in conditions:
        groups.setdefault(int(c["training_run"]), []).append(c)
    return [sorted(g, key=lambda c: int(c.get("perturb_length", 1))) for _, g in sorted(groups.items())]


def chain_cost(group, cost=None):
    """Expected epochs of a chained group: every member but the first skips the window of its predecessor."""
    cost = cost or expected_epochs
    total, prev = 0, None
    for c in group:
        total += cost(c) - (int(prev.get("perturb_length", 1)) if prev is not None else 0)
        prev = c
    return total


def _default_run_fn(config):
    from functions.new_cvpr_train_behavior_things_pipeline import run_behavioral_training
    return run_behavioral_training(config)


def cpu_threads_per_worker(n_workers):
    """Intra-op CPU threads of one sweep worker: the host's cores split over the workers, at most 4 (the GPU does
    the work; the host side of an epoch is Python), at least 1.
    Without a cap every worker process sizes its OpenMP / MKL pools for ALL cores; eight of them on one host
    (the full 136-condition grid on 8 B200s, profiles/r02_grid_full_136_n8.json) then spend their time spinning:
    0.85 s per epoch per worker against 0.086 s for one worker alone.  (torchrun ranks never see this: torchrun
    sets OMP_NUM_THREADS=1 itself.)"""
    try:
        cores = len(os.sched_getaffinity(0))
    except (AttributeError, OSError):
        cores = os.cpu_count() or 1
    try:   # a container's CPU quota (cgroup v2) is invisible to the affinity mask
        quota, period = open("/sys/fs/cgroup/cpu.max").read().split()[:2]
        if quota != "max":
            cores = min(cores, max(1, int(quota) // int(period)))
    except (OSError, ValueError):
        pass
    return max(1, min(4, cores // max(1, n_workers)))


_THREAD_ENV = ("OMP_NUM_THREADS", "MKL_NUM_THREADS", "OPENBLAS_NUM_THREADS", "NUMEXPR_NUM_THREADS")


def pinned_device_env(device_id, current=None):
    """The CUDA_VISIBLE_DEVICES value that pins a worker to the `device_id`-th device THIS process can see.
    `devices` entries of `run_sweep` are logical indices (what torch.cuda.device(i) would address here): when the
    launcher already restricted the job to some GPUs (CUDA_VISIBLE_DEVICES="4,5", a UUID list, a MIG slice), index i
    means the i-th entry of that list, not physical GPU i - which may belong to another job."""
    if current is None:
        current = os.environ.get("CUDA_VISIBLE_DEVICES")
    if current is None:
        return str(int(device_id))
    entries = [e.strip() for e in current.split(",") if e.strip()]
    if int(device_id) < 0 or int(device_id) >= len(entries):
        raise ValueError(f"run_sweep: device {device_id} requested but CUDA_VISIBLE_DEVICES={current!r} exposes "
                         f"{len(entries)} device(s)")
    return entries[int(device_id)]


def _worker(worker_id, device_id, tasks, results, base_config, layout, run_fn, chain=False, cpu_threads=0,
            device_env=None):
    # pin the GPU before anything initialises CUDA in this process
    if device_id is not None:
        os.environ["CUDA_VISIBLE_DEVICES"] = device_env if device_env is not None else pinned_device_env(device_id)
    if cpu_threads > 0:
        # (the environment set by run_sweep already sized the OpenMP / MKL pools at library load; this also covers
        # a parent that had the variables set to something else)
        try:
            import torch
            torch.set_num_threads(cpu_threads)
        except Exception:   # noqa: BLE001 - thread caps are an optimisation, never a reason to fail a condition
            pass
    run_fn = run_fn or _default_run_fn
    while True:
        item = tasks.get()
        if item is None:
            break
        for idx, cond in item:      # one condition, or one resume chain in increasing window length
            t0 = time.time()
            try:
                cfg = condition_config(base_config, cond, layout)
                if chain:
                    cfg["hba_resume_kind"] = apply_length_resume(cfg)
                if device_id is not None:
                    cfg["cuda"] = 0   # the only visible device (NEW:1137-1144)
                run_fn(cfg)
                results.put((idx, worker_id, True, time.time() - t0, ""))
            except Exception as exc:  # SWEEP:215-223: log, count, continue with the next condition
                results.put((idx, worker_id, False, time.time() - t0, f"{exc}\n{traceback.format_exc()}"))


def run_sweep(base_config, conditions, devices, layout="sweep", run_fn=None, cost=expected_epochs, log=print,
              chain=False):
    """Runs `conditions` on one worker process per entry of `devices` (CUDA device indices; None
    entries run without pinning, for CPU tests).  Returns a list of result dicts in condition order:
    {condition, worker, ok, seconds, error}.

    `chain=True` ('length' layout only) schedules LEN's resume chain (LEN:188-253): the conditions of one start
    epoch form one task, run on one worker in increasing window length, each resuming from the end of the
    previous, shorter window instead of recomputing it from the baseline epoch (`apply_length_resume`); the
    groups are handed out longest-first.  `chain=False`: every condition is its own task and resumes from
    the baseline checkpoint (independent conditions, best balance)."""
    import multiprocessing as mp
    if chain and layout != "length":
        raise ValueError("chain=True needs layout='length' (the single-epoch sweep has one window per start epoch)")
    ctx = mp.get_context("spawn")
    tasks, results = ctx.Queue(), ctx.Queue()
    index_of = {id(c): i for i, c in enumerate(conditions)}
    if chain:
        groups = chain_groups(list(conditions))
        for _, _, g in sorted(((-chain_cost(g, cost), i, g) for i, g in enumerate(groups))):
            tasks.put([(index_of[id(c)], c) for c in g])
    else:
        for c in lpt_order(list(conditions), cost):
            tasks.put([(index_of[id(c)], c)])
    for _ in devices:
        tasks.put(None)
    n_threads = cpu_threads_per_worker(len(devices))
    # (resolved here, against the launcher's own CUDA_VISIBLE_DEVICES: a bad index fails before anything is spawned)
    pins = [None if dev is None else pinned_device_env(dev) for dev in devices]
    procs = [ctx.Process(target=_worker, args=(w, dev, tasks, results, base_config, layout, run_fn, chain, n_threads, pin),
                         daemon=False) for w, (dev, pin) in enumerate(zip(devices, pins))]
    t0 = time.time()
    # the children read these when their OpenMP / MKL runtimes load (spawn: fresh interpreters inheriting os.environ);
    # passive waiting keeps idle pool threads off the cores the other workers' Python threads need
    saved_env = {k: os.environ.get(k) for k in _THREAD_ENV + ("OMP_WAIT_POLICY",)}
    for k in _THREAD_ENV:
        os.environ[k] = str(n_threads)
    os.environ["OMP_WAIT_POLICY"] = "passive"
    try:
        for p in procs:
            p.start()
    finally:
        for k, v in saved_env.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
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
