"""Dry-runs the GPU test files on the CPU (no GPU in the authoring container) with
libhba served by its CPU restatement (oracle/libhba_ref.py) - the test functions themselves, through pytest, with
`DEV` pointed at the CPU and CUDA-graph capture off.  NOT a test and NOT a product path: it catches host-side mistakes
(names, shapes, state handling, file formats, cross-condition state) in code that otherwise only executes on a B200.
Tests that need the device itself (CUDA generators, device-side timing) fail or are meaningless here; captured
graphs are replaced by call-by-call launches.

    python tests/emulate_clip_gpu_tests.py tests/test_gpu_pipeline.py [-k expr]

tests/test_host_on_ref_lib_cpu.py runs it over DRY_RUN_FILES as part of the CPU suite: besides the host logic of the
product this pins the restatement itself - under it, tests/test_gpu_ops.py compares every restated entry point with the
very formulas (torch / fp64 autograd / numpy / scipy) that judge the CUDA kernels on a B200.
"""
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "vit-project_b200"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


class _Plugin:
    def __init__(self):
        self.ctx = None

    def pytest_sessionstart(self, session):
        from oracle.libhba_ref import emulated_device
        import functions._pipeline_core as core
        import functions.cvpr_train_behavior_things_pipeline_baseline as BASE
        import functions.new_cvpr_train_behavior_things_pipeline as NEW
        self.ctx = emulated_device()
        self.ctx.__enter__()
        cpu = lambda flag: torch.device("cpu")
        core.select_device = BASE.select_device = NEW.select_device = cpu
        # no CUDA-graph capture without a device: every step is launched call by call, whatever a test asks for
        # (a "captured = eager" comparison then compares eager with eager; the tests AROUND the graphs - the fused
        # MSE step against the criterion-called step, NaN batches, evaluation - keep their meaning)
        from hba import vit
        core.TrainStep._graphs_on = lambda self: False
        core._CachedForwardGraphs.usable = lambda self, loader: False
        vit.DataParallelTrainer.step = lambda self, images, labels: self._step_eager(images, labels)

    def pytest_sessionfinish(self, session, exitstatus):
        if self.ctx is not None:
            self.ctx.__exit__(None, None, None)

    @pytest.hookimpl(trylast=True)
    def pytest_collection_modifyitems(self, config, items):
        for item in items:
            item.own_markers = [m for m in item.own_markers if not (m.name == "skip" and "CUDA" in str(m.kwargs))]
            if hasattr(item.module, "DEV"):
                item.module.DEV = torch.device("cpu")


# what cannot run without the device itself: one comparison whose 1e-6 tolerance is about the summation order of two
# CUDA kernels (fused MSE head vs head + torch MSE), which the restatement does not reproduce (captured CUDA graphs are
# replaced by call-by-call launches, see the plugin; the front ends' argument checks are restated)
NEEDS_DEVICE = "not fused_mse_step_equals"
DRY_RUN_FILES = ("test_gpu_ops.py", "test_gpu_model.py", "test_gpu_pipeline.py", "test_gpu_vit.py",
                 "test_gpu_zzz_general_placement.py")


if __name__ == "__main__":
    args = sys.argv[1:]
    if "-k" not in args:
        args += ["-k", NEEDS_DEVICE]
    sys.exit(pytest.main(args + ["-q", "-p", "no:cacheprovider", "-m", "gpu"], plugins=[_Plugin()]))
