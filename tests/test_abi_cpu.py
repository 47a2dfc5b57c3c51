"""CPU checks of the C-ABI boundary: the library loads, exports every symbol include/hba.h declares,
the ctypes table agrees with the header's prototypes, and the product path refuses to run without a
CUDA device (no CPU fallback)."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "hba.h")


def header_prototypes():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    src = re.sub(r"typedef struct.*?hba_gemm_params;", "", src, flags=re.S)
    src = re.sub(r"enum\s*\{.*?\};", "", src, flags=re.S)
    protos = {}
    for m in re.finditer(r"([\w\s\*]+?)\b(hba_\w+)\s*\(([^;{]*)\)\s*;", src):
        args = m.group(3).strip()
        n = 0 if args in ("", "void") else len(args.split(","))
        protos[m.group(2)] = n
    return protos


def test_header_declares_expected_entry_points():
    protos = header_prototypes()
    assert len(protos) >= 26
    for name in ("hba_gemm_bf16", "hba_dora_merge_fwd", "hba_dora_merge_bwd", "hba_cos_head_fwd",
                 "hba_rank_avg_f64", "hba_rdm_f64", "hba_adamw_multi", "hba_attention_fwd"):
        assert name in protos


def test_library_exports_every_declared_symbol_and_ctypes_table_matches():
    from hba import _lib
    lib = _lib.load()
    protos = header_prototypes()
    assert set(protos) == set(_lib.SIGNATURES), set(protos) ^ set(_lib.SIGNATURES)
    for name, nargs in protos.items():
        assert hasattr(lib, name), f"libhba.so does not export {name}"
        assert len(_lib.SIGNATURES[name][1]) == nargs, (name, nargs, len(_lib.SIGNATURES[name][1]))
    assert lib.hba_abi_version() == 1
    assert isinstance(_lib.last_error(), str)


def test_gemm_params_struct_layout_matches_header():
    from hba import _lib
    src = open(HEADER).read()
    body = re.search(r"typedef struct hba_gemm_params \{(.*?)\} hba_gemm_params;", src, re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    names = []
    for decl in body.split(";"):
        decl = decl.strip()
        if not decl:
            continue
        decl = re.sub(r"^(const\s+)?(void|float|int32_t)\s*\**", "", decl)
        names += [n.strip().lstrip("*") for n in decl.split(",")]
    assert names == [f[0] for f in _lib.GemmParams._fields_]


def test_argument_errors_are_reported_not_fatal():
    from hba import _lib
    lib = _lib.load()
    rc = lib.hba_gemm_bf16(None, None)
    assert rc == -22 and "null" in _lib.last_error()
    rc = lib.hba_rank_avg_f64(None, 0, None, None, 0, None)
    assert rc == -22


def test_no_cpu_fallback():
    from src.models.CLIPs.clip_hba import clip
    model = clip.build_model(clip.synthetic_state_dict("ViT-tiny-14.pt"))
    img = torch.zeros(1, 3, 224, 224)
    tok = clip.tokenize(["a", "b"])
    with pytest.raises(RuntimeError, match="no CPU path"):
        model(img, tok, True)
    with pytest.raises(NotImplementedError):
        model.visual(img)
    import hba
    layer = hba.DoRALayer(torch.nn.Linear(64, 64), r=8)
    with pytest.raises(RuntimeError, match="no CPU path"):
        _ = layer.weight


def test_plugin_module_surface():
    """the four attributes the reference uses (NEW:251-265, 282) + module paths it patches."""
    from src.models.CLIPs.clip_hba import clip
    assert "ViT-L/14" in clip._MODELS
    t = clip.tokenize("metallic; artificial")
    assert t.shape == (1, 77) and t.dtype == torch.long and int(t.argmax()) == 3
    model = clip.build_model(clip.synthetic_state_dict("ViT-tiny-14.pt"))
    blk = model.visual.transformer.resblocks[-1]
    assert isinstance(blk.attn, torch.nn.MultiheadAttention)
    assert blk.attn.out_proj.weight.shape == (256, 256)
    assert model.transformer.resblocks[-1].attn.out_proj.in_features == 128
    assert not model.training
