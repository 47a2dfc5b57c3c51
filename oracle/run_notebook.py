"""ORACLE (test infrastructure only - never imported by the product path).

Executes the code cells of one of the reference's figure notebooks (Figures/fig*/fig*.ipynb), unmodified, in one
namespace with the notebook's directory as working directory (their data paths are relative: '../../Data/...'),
and returns that namespace.  matplotlib / seaborn are not installed here and are irrelevant to the numbers: they
are replaced by mocks that accept every call (and unpack as pairs, for `fig, ax = plt.subplots()` /
`xmin, xmax = ax.get_xlim()`).  Used by oracle/make_analysis_golden.py and tests/test_analysis_cpu.py to compare
the analysis layer (hba.analysis) with the tables the notebooks themselves compute.
"""
import contextlib
import io
import json
import os
import sys

FIGURES = "/root/reference/Figures"
NOTEBOOKS = {"fig2": "fig2 (Effects of Different Perturbations)/fig2.ipynb",
             "fig3": "fig3 (Single Sweep Perturbation Experiments)/fig3.ipynb",
             "fig4": "fig4 (Perturbation Recovery)/fig4.ipynb"}


class _Pair:
    """Stands for any plotting object: every attribute, call, item and arithmetic result is another one; it
    unpacks as a pair (`fig, ax = plt.subplots()`, `xmin, xmax = ax.get_xlim()`) and converts to the number 1."""

    def __getattr__(self, name):
        if name.startswith("__") and name.endswith("__"):
            raise AttributeError(name)
        return _Pair()

    def __call__(self, *a, **k):
        return _Pair()

    def __iter__(self):
        return iter((_Pair(), _Pair()))

    def __getitem__(self, key):
        return _Pair()

    def __setitem__(self, key, value):
        pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False

    def __int__(self):
        return 1

    def __float__(self):
        return 1.0

    def __index__(self):
        return 1

    def __bool__(self):
        return True

    def _num(self, *_):
        return 1.0

    __add__ = __radd__ = __sub__ = __rsub__ = __mul__ = __rmul__ = __truediv__ = __rtruediv__ = _num
    __floordiv__ = __rfloordiv__ = __neg__ = __abs__ = __mod__ = _num


def run_cells(name, stop_after=None):
    """-> namespace after executing the notebook's code cells (up to and including cell index `stop_after`)."""
    path = os.path.join(FIGURES, NOTEBOOKS[name])
    cells = json.load(open(path))["cells"]
    import numpy  # noqa: F401  (imported before the stubs go in: the cells must find the real modules loaded)
    import pandas  # noqa: F401
    stubs = {m: _Pair() for m in ("matplotlib", "matplotlib.pyplot", "matplotlib.patches", "matplotlib.colors",
                                  "matplotlib.ticker", "matplotlib.lines", "seaborn")}
    saved = {m: sys.modules.get(m) for m in stubs}
    ns = {"__name__": "__notebook__"}
    cwd = os.getcwd()
    os.chdir(os.path.dirname(path))
    sys.modules.update(stubs)
    try:
        with contextlib.redirect_stdout(io.StringIO()):
            for i, c in enumerate(cells):
                if c["cell_type"] != "code":
                    continue
                src = "".join(c["source"])
                src = "\n".join(l for l in src.splitlines() if not l.lstrip().startswith(("%", "!")))
                exec(compile(src, f"{name}.ipynb[cell {i}]", "exec"), ns)
                if stop_after is not None and i >= stop_after:
                    break
    finally:
        os.chdir(cwd)
        for m, old in saved.items():
            if old is None:
                sys.modules.pop(m, None)
            else:
                sys.modules[m] = old
    return ns
