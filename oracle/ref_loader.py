"""ORACLE (test infrastructure only).

Imports the reference's OWN pipeline modules, read-only, from /root/reference/Training with
the un-vendored ``src.models.CLIPs.clip_hba.clip`` dependency replaced by the restated
``oracle.clip_ref``.  Only usable in the authoring container (the GPU box has no
/root/reference); it is used by ``oracle/make_golden.py`` to generate the committed fixtures
under tests/golden/ and by the CPU tests that pin the restatement (they skip when the
reference is absent).

Reference modules made importable:
  NEW  = Training/functions/new_cvpr_train_behavior_things_pipeline.py
  BASE = Training/functions/cvpr_train_behavior_things_pipeline_baseline.py
"""
from __future__ import annotations

import importlib
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("HBA_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "Training", "functions"))


def _save_dora_parameters_stub(model, path, epoch, vision_layers, transformer_layers, log_fn=None):
    """Stand-in for the un-vendored ``src.models.clip_hba_utils.save_dora_parameters``
    (BASE:22, BASE:683-690): same on-disk format as NEW:657-693."""
    import torch
    mod = model.module if isinstance(model, torch.nn.DataParallel) else model
    out = {}
    vb = mod.clip_model.visual.transformer.resblocks
    tb = mod.clip_model.transformer.resblocks
    paths = [f"clip_model.visual.transformer.resblocks.{len(vb) - vision_layers + i}.attn.out_proj"
             for i in range(vision_layers)]
    paths += [f"clip_model.transformer.resblocks.{len(tb) - transformer_layers + i}.attn.out_proj"
              for i in range(transformer_layers)]
    for p in paths:
        m = mod
        for a in p.split("."):
            m = getattr(m, a)
        for n in ("m", "delta_D_A", "delta_D_B"):
            out[f"{p}.{n}"] = getattr(m, n).detach().cpu()
    os.makedirs(path, exist_ok=True)
    torch.save(out, os.path.join(path, f"epoch{epoch + 1}_dora_params.pth"))


def load_reference(clip_module=None):
    """Returns (NEW, BASE) reference modules.  ``clip_module`` defaults to oracle.clip_ref."""
    if not reference_available():
        raise RuntimeError(f"reference not found under {REFERENCE_ROOT}")
    if clip_module is None:
        from oracle import clip_ref as clip_module
    # the stubs below live in sys.modules only while the reference is imported (the product ships
    # its own `src.models...` plug-in package under the same names)
    saved_src = {k: sys.modules.pop(k) for k in list(sys.modules) if k == "src" or k.startswith("src.")}
    for name in ("src", "src.models", "src.models.CLIPs", "src.models.CLIPs.clip_hba"):
        if name not in sys.modules or not getattr(sys.modules[name], "__hba_stub__", False):
            m = types.ModuleType(name)
            m.__path__ = []
            m.__hba_stub__ = True
            sys.modules[name] = m
    sys.modules["src.models.CLIPs.clip_hba.clip"] = clip_module
    sys.modules["src.models.CLIPs.clip_hba"].clip = clip_module
    utils = types.ModuleType("src.models.clip_hba_utils")
    utils.save_dora_parameters = _save_dora_parameters_stub
    utils.__hba_stub__ = True
    sys.modules["src.models.clip_hba_utils"] = utils
    tr = os.path.join(REFERENCE_ROOT, "Training")
    # the reference must win the name "functions" while it is being imported
    saved = {k: sys.modules.pop(k) for k in list(sys.modules)
             if k == "functions" or k.startswith("functions.")}
    # a regular package named `functions` anywhere on sys.path (the product's drop-in package) would
    # shadow the reference's namespace package: hide such entries while the reference is imported
    hidden = [p for p in sys.path if os.path.exists(os.path.join(p or ".", "functions", "__init__.py"))]
    for p in hidden:
        sys.path.remove(p)
    sys.path.insert(0, tr)
    try:
        new = importlib.import_module("functions.new_cvpr_train_behavior_things_pipeline")
        base = importlib.import_module("functions.cvpr_train_behavior_things_pipeline_baseline")
    finally:
        sys.path.remove(tr)
        sys.path[:0] = hidden
        for k in [k for k in sys.modules if k == "functions" or k.startswith("functions.")]:
            sys.modules["_reference_" + k] = sys.modules.pop(k)
        sys.modules.update(saved)
        for k in [k for k in sys.modules if k == "src" or k.startswith("src.")]:
            sys.modules.pop(k)
        sys.modules.update(saved_src)
    return new, base
