"""Development aid (no GPU in the authoring container): dry-runs tests/test_gpu_vit_measure.py on the CPU with
the device layer of hba.vit emulated in torch, to catch host-side mistakes (names, shapes, state handling,
file formats) in code that otherwise only executes on a B200.  NOT a test and NOT a product path:

  * `ViTEngine.forward` / `.backward` are replaced by the oracle model + torch autograd,
  * `ops.softmax_ce`, `ops.sgd_multi`, `ops.sgd_staged` by torch / ctypes equivalents on the same buffers,
  * `hba.rsa.RSAEvaluator` by the NumPy / SciPy tail,
  * the `is_cuda` guards are stripped and "cuda" device strings mapped to the CPU.

    python tests/emulate_vit_measure_tests.py [-k substring]

Numerical tolerances of the real tests are meaningless here (the emulation IS the oracle); a run that ends
with "all emulated tests ran" only says the Python around the kernels is coherent.
"""
import ctypes
import inspect
import os
import sys
import tempfile
import textwrap
from pathlib import Path

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "vit-project_b200"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

from hba import ops, rsa, vit  # noqa: E402
from oracle import vit_measure_ref, vit_ref  # noqa: E402


def _oracle_of(model):
    cfg = dict(embed_dim=model.embed_dim, depth=len(model.blocks), num_heads=model.blocks[0].attn.num_heads,
               num_classes=model.num_classes, img_size=model.img_size, patch_size=model.patch_size)
    o = vit_ref.VisionTransformerRef(**cfg)
    o.load_state_dict(model.state_dict())
    return o


def emu_forward(self, images, save=True, features=False):
    self._setup(images.device)
    o = _oracle_of(self.m)
    if features:
        with torch.no_grad():
            return o.forward_features(images.float())
    if save:
        logits = o(images.float())
        self._emu = (o, logits)
        self._saved_B = images.shape[0]
        out = self._buf("logits", (images.shape[0], (self.m.num_classes + 3) // 4 * 4))[:, :self.m.num_classes]
        out.copy_(logits.detach())
        return out
    with torch.no_grad():
        out = self._buf("logits", (images.shape[0], (self.m.num_classes + 3) // 4 * 4))[:, :self.m.num_classes]
        out.copy_(o(images.float()))
        return out


def emu_backward(self, d_logits, on_bucket_ready=None):
    self._ensure_grads()
    o, logits = self._emu
    names = [n for n, _ in self.m.named_parameters()]
    grads = torch.autograd.grad(logits, [dict(o.named_parameters())[n] for n in names], d_logits.clone())
    for (n, p), g in zip(self.m.named_parameters(), grads):
        self.grad_of[id(p)].copy_(g)
    if on_bucket_ready is not None:
        for name, s, e in self.bucket_slices:
            on_bucket_ready(name, self.flat_grad[s:e])
    return self.flat_grad


def emu_softmax_ce(logits, labels, loss, d_logits, correct, workspace):
    B = logits.shape[0]
    loss.copy_(F.cross_entropy(logits, labels).reshape(1))
    if d_logits is not None:
        d_logits.copy_((torch.softmax(logits, 1) - F.one_hot(labels, logits.shape[1]).float()) / B)
    if correct is not None:
        correct.copy_(logits.max(1)[1].eq(labels).sum().to(torch.int32).reshape(1))


def _arr(ptr, n):
    return np.ctypeslib.as_array((ctypes.c_float * n).from_address(ptr))


def _sgd(rows, lr, momentum, wd, first):
    for p_ptr, g_ptr, m_ptr, n in rows:
        p, g, m = _arr(p_ptr, n), _arr(g_ptr, n), _arr(m_ptr, n)
        gg = g + np.float32(wd) * p
        m[:] = gg if first else np.float32(momentum) * m + gg
        p -= np.float32(lr) * m


def emu_sgd_multi(table, sizes, n, total, lr, momentum, wd, first_step, skip_flag=None):
    t, s = table.tolist(), sizes.tolist()
    _sgd([(t[3 * i], t[3 * i + 1], t[3 * i + 2], s[i]) for i in range(n)], lr, momentum, wd, first_step)


def emu_sgd_staged(table4, prefix4, n, total4, lr, momentum, wd, first_step, skip_flag=None):
    t, pre = table4.tolist(), prefix4.tolist() + [total4]
    _sgd([(t[4 * i], t[4 * i + 1], t[4 * i + 2], 4 * (pre[i + 1] - pre[i])) for i in range(n)], lr, momentum, wd,
         first_step)


class EmuRSA:
    def __init__(self, reference_rdm, device="cpu"):
        self.ref, self.N = np.asarray(reference_rdm, dtype=np.float64), len(reference_rdm)
        self.P = self.N * (self.N - 1) // 2

    def __call__(self, emb, want_rdm=True):
        rho, p = vit_measure_ref.rsa_tail_ref(emb.detach().cpu().numpy(), self.ref)
        return float(rho), float(p), None


def _strip_guard(fn):
    src = textwrap.dedent(inspect.getsource(fn))
    lines, out, skip = src.splitlines(), [], 0
    for ln in lines:
        if skip:
            skip -= 1
            continue
        if "if not x.is_cuda" in ln:
            skip = 1
            continue
        out.append(ln)
    ns = {}
    exec(compile("\n".join(out), f"<emulated {fn.__name__}>", "exec"), vit.__dict__, ns)
    return ns[fn.__name__]


class _TorchProxy:
    """`torch` as seen by the scripts: cuda devices become the CPU."""

    def __init__(self):
        self.cuda = type("cuda", (), {"set_device": staticmethod(lambda *_: None),
                                      "current_device": staticmethod(lambda: 0)})

    def device(self, *a, **k):
        return torch.device("cpu")

    def __getattr__(self, name):
        return getattr(torch, name)


def install():
    vit.ViTEngine.forward = emu_forward
    vit.ViTEngine.backward = emu_backward
    vit.VisionTransformer.forward_features = _strip_guard(vit.VisionTransformer.forward_features)
    vit.VisionTransformer.forward = _strip_guard(vit.VisionTransformer.forward)
    ops.softmax_ce, ops.sgd_multi, ops.sgd_staged = emu_softmax_ce, emu_sgd_multi, emu_sgd_staged
    rsa.RSAEvaluator = EmuRSA
    from hba import vit_train
    vit_train.setup_distributed = lambda: (print("Not using distributed mode (emulated)"), (0, 1, 0))[1]
    real_trainer_init = vit.DataParallelTrainer.__init__

    def init(self, *a, **k):
        k["use_graph"] = False                      # no CUDA graphs on the CPU
        real_trainer_init(self, *a, **k)
    vit.DataParallelTrainer.__init__ = init

    import contextlib
    import pytest
    import test_gpu_vit_measure as t
    t.DEV = "cpu"

    class _Pytest:
        """pytest as seen by the tests: the "no CPU path" RuntimeError checks do not apply under emulation."""

        def __getattr__(self, name):
            return getattr(pytest, name)

        @staticmethod
        def raises(exc, *a, **k):
            if exc is RuntimeError:
                return contextlib.suppress(RuntimeError)
            return pytest.raises(exc, *a, **k)
    t.pytest = _Pytest()
    real_load = t._load_script

    def load_script(rel):
        mod = real_load(rel)
        mod.torch = _TorchProxy()
        return mod
    t._load_script = load_script
    return t


def main():
    key = sys.argv[sys.argv.index("-k") + 1] if "-k" in sys.argv else ""
    t = install()
    ran = []
    for name, fn in sorted(vars(t).items()):
        if not name.startswith("test_") or key not in name:
            continue
        marks = [m for m in getattr(fn, "pytestmark", []) if m.name == "parametrize"]
        cases = [()]
        if marks:
            argnames = [a.strip() for a in marks[0].args[0].split(",")]
            cases = [c if isinstance(c, tuple) else (c,) for c in marks[0].args[1]]
        for case in cases:
            kwargs = dict(zip(argnames, case)) if marks else {}
            if "tmp_path" in inspect.signature(fn).parameters:
                kwargs["tmp_path"] = Path(tempfile.mkdtemp(prefix="emu_"))
            print(f"--- {name} {kwargs if marks else ''}", flush=True)
            fn(**kwargs)
            ran.append(name)
    print(f"all emulated tests ran: {len(ran)} cases")


if __name__ == "__main__":
    main()
