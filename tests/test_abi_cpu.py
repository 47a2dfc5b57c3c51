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


def _write_vocab(path, merges):
    import gzip
    with gzip.open(path, "wt", encoding="utf-8") as f:
        f.write("#version: test\n" + "\n".join(merges) + "\n")


def test_bpe_tokenizer_merges_in_rank_order(tmp_path, monkeypatch):
    """simple_tokenizer.py on a hand-made merge table: ids = 512 + merge index for merged symbols, byte symbols
    in the published order ('!' = 0 ... , the same with </w> from 256), SOT / EOT last."""
    from src.models.CLIPs.clip_hba import simple_tokenizer as st
    vocab = tmp_path / "bpe_simple_vocab_16e6.txt.gz"
    _write_vocab(vocab, ["h e", "he l", "l o</w>", "w o", "r l", "rl d</w>"])
    monkeypatch.setenv("HBA_BPE_VOCAB", str(vocab))
    st.load.cache_clear()
    tok = st.load()
    try:
        sym = lambda ch: ord(ch) - ord("!")
        # hello -> h e l l o</w> -> he l l o</w> -> hel l o</w> -> hel lo</w>
        assert tok.encode("hello") == [512 + 1, 512 + 2]
        # "World!" is lower-cased; wo | rl d</w> -> wo rld</w>;  "!" is its own token with </w>
        assert tok.encode("  World! ") == [512 + 3, 512 + 5, 256 + sym("!")]
        assert tok.encode("a&amp;b") == [256 + sym("a"), 256 + sym("&"), 256 + sym("b")]   # html unescape
        assert tok.sot == tok.eot - 1 == len(tok.encoder) - 2
    finally:
        st.load.cache_clear()


def test_tokenize_refuses_made_up_ids_for_a_real_checkpoint(tmp_path, monkeypatch):
    """the word-hash pseudo ids are for synthetic (random-init) checkpoints only."""
    from src.models.CLIPs.clip_hba import clip, simple_tokenizer as st
    monkeypatch.delenv("HBA_TOKENIZER", raising=False)
    monkeypatch.setenv("HBA_BPE_VOCAB", str(tmp_path / "missing.gz"))
    monkeypatch.setattr(st, "vocab_candidates", lambda: [str(tmp_path / "missing.gz")])
    st.load.cache_clear()
    real = tmp_path / "ViT-L-14.pt"
    real.write_bytes(b"x")                              # a cached "published" checkpoint
    saved = dict(clip._CHECKPOINT)
    try:
        assert clip._download(clip._MODELS["ViT-L/14"], str(tmp_path)) == str(real)
        assert clip._CHECKPOINT["synthetic"] is False
        with pytest.raises(RuntimeError, match="BPE merge table"):
            clip.tokenize("metallic; artificial")
        monkeypatch.setenv("HBA_TOKENIZER", "pseudo")
        with pytest.warns(RuntimeWarning, match="pseudo ids"):
            t = clip.tokenize("metallic; artificial")
        assert t.shape == (1, 77)
        monkeypatch.delenv("HBA_TOKENIZER")
        # with a merge table the published encoding is used
        vocab = tmp_path / "bpe_simple_vocab_16e6.txt.gz"
        _write_vocab(vocab, ["h e", "he l", "l o</w>"])
        monkeypatch.setattr(st, "vocab_candidates", lambda: [str(vocab)])
        st.load.cache_clear()
        t = clip.tokenize("hello")
        tok = st.load()
        assert t[0, :4].tolist() == [tok.sot, 513, 514, tok.eot] and int(t[0].argmax()) == 3
        # a synthetic checkpoint goes back to the word-hash ids (what the oracle / goldens use)
        clip._download("synthetic://ViT-tiny-14.pt", str(tmp_path))
        assert clip._CHECKPOINT["synthetic"] is True
        assert clip.tokenize("hello")[0, 0] == clip.SOT_TOKEN and int(clip.tokenize("hello")[0, 2]) == clip.EOT_TOKEN
    finally:
        clip._CHECKPOINT.update(saved)
        st.load.cache_clear()


def _download_worker(root, out):
    import time
    import torch
    from src.models.CLIPs.clip_hba import clip
    t0 = time.time()
    path = clip._download(clip._MODELS["ViT-tiny/14"], root)
    sd = torch.load(path, map_location="cpu")          # must never be a half-written file
    with open(out, "w") as f:
        f.write(f"{os.getpid()} {os.path.getmtime(path)!r} {len(sd)} {float(sd['logit_scale'])!r} {t0!r}")


def test_synthetic_checkpoint_is_materialised_once_under_concurrent_first_use(tmp_path):
    """Eight sweep workers / torchrun ranks on a fresh machine all find no cached checkpoint at the same moment: one
    generates it under a lock, the file appears atomically, nobody loads a partial file, and a completed file is
    never replaced (its mtime keys the per-process frozen-model cache)."""
    import multiprocessing as mp
    root = str(tmp_path / "clip_cache")
    ctx = mp.get_context("spawn")
    outs = [str(tmp_path / f"w{i}.txt") for i in range(6)]
    procs = [ctx.Process(target=_download_worker, args=(root, o)) for o in outs]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
        assert p.exitcode == 0
    rows = [open(o).read().split() for o in outs]
    assert len({r[0] for r in rows}) == 6                      # six processes ...
    assert len({r[1] for r in rows}) == 1                      # ... saw one and the same file (mtime unchanged)
    assert len({(r[2], r[3]) for r in rows}) == 1
    left = sorted(os.listdir(root))
    assert left == ["ViT-tiny-14.pt", "ViT-tiny-14.pt.lock"], left      # no temporary files stay behind
