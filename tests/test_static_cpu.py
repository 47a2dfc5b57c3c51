"""Static coherence of the Python that only executes on a B200 (engine, CUDA-graph paths, bench sections, tools):
there is no GPU in the authoring container, so a misspelt name in such a path would surface at round end only.
Two stdlib-only checks over every product / bench / tool file: (1) every global name a function references is
defined at module level, imported or a builtin (symtable); (2) every `alias.attr` where `alias` is an imported
module resolves on the imported module (ast + importlib; the modules import on the CPU)."""
import ast
import builtins
import importlib
import os
import symtable
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "vit-project_b200")
TARGETS = [os.path.join(ROOT, "bench.py"), os.path.join(ROOT, "__graft_entry__.py"), PKG, os.path.join(ROOT, "tools"),
           os.path.join(ROOT, "oracle")]
IMPLICIT = {"__file__", "__name__", "__doc__", "__builtins__", "__spec__", "__package__"}


def _files():
    for t in TARGETS:
        if t.endswith(".py"):
            yield t
            continue
        for d, _, fs in os.walk(t):
            for f in sorted(fs):
                if f.endswith(".py"):
                    yield os.path.join(d, f)


def test_no_undefined_global_names():
    suspects = []

    def walk(tab, module_names, path):
        for child in tab.get_children():
            for s in child.get_symbols():
                n = s.get_name()
                if (s.is_referenced() and s.is_global() and not s.is_assigned() and n not in module_names
                        and not hasattr(builtins, n)):
                    suspects.append(f"{os.path.relpath(path, ROOT)}:{child.get_lineno()} {child.get_name()}(): {n}")
            walk(child, module_names, path)
    n_files = 0
    for path in _files():
        src = open(path).read()
        n_files += 1
        tab = symtable.symtable(src, path, "exec")
        names = {s.get_name() for s in tab.get_symbols() if s.is_assigned() or s.is_imported() or s.is_namespace()}
        names |= IMPLICIT
        if "import *" in src:      # (the one star re-export module: names come from the star)
            continue
        for s in tab.get_symbols():
            n = s.get_name()
            if s.is_referenced() and n not in names and not hasattr(builtins, n):
                suspects.append(f"{os.path.relpath(path, ROOT)}: <module>: {n}")
        walk(tab, names, path)
    assert n_files > 40 and not suspects, "\n".join(suspects)


def _package_of(path):
    rel = os.path.relpath(path, PKG)
    if rel.startswith(".."):
        return None
    return ".".join(rel[:-3].split(os.sep)[:-1])


def test_module_attribute_references_resolve():
    os.environ.setdefault("HBA_SYNTHETIC_OK", "1")
    suspects, checked = [], 0
    for path in _files():
        if os.sep + "oracle" + os.sep in path:     # (oracle modules juggle sys.modules to import the reference)
            continue
        tree = ast.parse(open(path).read())
        pkg = _package_of(path)
        alias, local_rebinds = {}, set()
        for node in ast.walk(tree):
            if isinstance(node, ast.Import):
                for a in node.names:
                    try:
                        if a.asname:
                            alias[a.asname] = importlib.import_module(a.name)
                        elif "." not in a.name:
                            alias[a.name] = importlib.import_module(a.name)
                    except ImportError:
                        pass       # optional third-party modules (timm, torchvision in a stubbed environment)
            elif isinstance(node, ast.ImportFrom):
                try:
                    if node.level:
                        if pkg is None:
                            continue
                        mod = importlib.import_module("." * node.level + (node.module or ""), package=pkg)
                    else:
                        mod = importlib.import_module(node.module)
                except ImportError:
                    continue
                for a in node.names:
                    if a.name == "*":
                        continue
                    obj = getattr(mod, a.name, None)
                    if obj is None:
                        try:
                            obj = importlib.import_module(mod.__name__ + "." + a.name)
                        except ImportError:
                            suspects.append(f"{os.path.relpath(path, ROOT)}:{node.lineno} from {mod.__name__} import {a.name}")
                            continue
                    if isinstance(obj, types.ModuleType):
                        alias[a.asname or a.name] = obj
            elif isinstance(node, (ast.Assign, ast.AnnAssign, ast.AugAssign, ast.For, ast.With, ast.arg)):
                for t in ast.walk(node) if not isinstance(node, ast.arg) else [node]:
                    if isinstance(t, ast.Name) and isinstance(t.ctx, ast.Store):
                        local_rebinds.add(t.id)
                    elif isinstance(t, ast.arg):
                        local_rebinds.add(t.arg)
        for node in ast.walk(tree):
            if (isinstance(node, ast.Attribute) and isinstance(node.value, ast.Name) and node.value.id in alias
                    and node.value.id not in local_rebinds):
                mod = alias[node.value.id]
                checked += 1
                if not hasattr(mod, node.attr):
                    try:
                        importlib.import_module(mod.__name__ + "." + node.attr)
                    except ImportError:
                        suspects.append(f"{os.path.relpath(path, ROOT)}:{node.lineno} {node.value.id}.{node.attr} "
                                        f"(module {mod.__name__})")
    assert checked > 1000 and not suspects, "\n".join(suspects)
