"""Numeric parity at BASELINE.json's sizes.  The checker is the oracle (oracle/clip_ref.py + dora_ref.py,
oracle/vit_ref.py, numpy / scipy) run ON THE SAME GPU in torch fp32 with TF32 switched off - the same
arithmetic the CPU oracle performs, minutes faster; it is the checker only, never the thing measured.

  * CLIP-HBA, ViT-L/14 + DoRA r32 (last 2 vision + 1 text block), batch 32, the 66 SPoSE prompts
    (reference: CLIPHBA.forward NEW:287-304, DoRALayer.weight NEW:447-463, MSE NEW:994):
    predictions / loss / the 9 DoRA gradients - fp32 mode within 1e-3 relative (north star), bf16 mode
    within the tolerances stated in BF16_TOL below;
  * a 24-step bf16-mode training trajectory (AdamW 3e-4, NEW:1181) against the fp32 oracle: per-step loss and
    behavioural-RSA rho (NEW:605-654) on 48 held-out images every 8 steps;
  * ViT-B/16 at batch 256 (VIT:125-165): logits / loss / every parameter gradient, fp32 and bf16 modes;
  * RSA at scale (config 5): RDM, average-tie ranks and Spearman rho of 1,854 x 66 embeddings against
    numpy.corrcoef / scipy.stats.rankdata / scipy.stats.spearmanr (rho 1e-10, ranks bit-exact).

Every measured error is appended to gpurun_out/parity_fullsize.json so that the tolerances quoted in
DESIGN.md section 4 can be read off a run."""
import json
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda:0")

# stated bf16-mode tolerances at full size (24 + 12 blocks of bf16 operands, fp32 accumulate)
# (set from the first B200 run, profiles/r02_parity_fullsize.json: 7.0e-3 / 2.6e-4 / 8.9e-3 / 0.99998 / 1.7e-4 / 7.9e-4,
# with a margin of ~3x)
BF16_TOL = {"pred_rel_max": 2.5e-2,    # max |pred - oracle| / max |oracle|
            "loss_rel": 2e-3,
            "grad_rel_l2": 3e-2,       # ||g - g_oracle||_2 / ||g_oracle||_2 per DoRA tensor
            "grad_cos": 0.999,         # cosine(g, g_oracle) per DoRA tensor
            "traj_loss_rel": 2e-3,     # every step of the 24-step trajectory
            "traj_rho_abs": 5e-3}      # behavioural-RSA rho along the trajectory


def _record(name, values):
    path = os.path.join(ROOT, "gpurun_out", "parity_fullsize.json")
    try:
        os.makedirs(os.path.dirname(path), exist_ok=True)
        data = json.load(open(path)) if os.path.exists(path) else {}
        data[name] = values
        json.dump(data, open(path, "w"), indent=1)
    except OSError:
        pass


def _exact_fp32():
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    torch.set_float32_matmul_precision("highest")


def rel_max(a, b):
    a, b = a.detach().double(), b.detach().double()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def rel_l2(a, b):
    a, b = a.detach().double(), b.detach().double()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def cosine(a, b):
    a, b = a.detach().double().flatten(), b.detach().double().flatten()
    return float((a @ b) / (a.norm() * b.norm()).clamp_min(1e-30))


# ------------------------------------------------------------------------------ CLIP-HBA, ViT-L/14
@pytest.fixture(scope="module")
def clip_state():
    from oracle import clip_ref
    from functions.spose_dimensions import classnames66
    sd = clip_ref.synthetic_state_dict("ViT-L/14", seed=1)
    tokens = torch.stack([clip_ref.tokenize(c) for c in classnames66])
    g = torch.Generator().manual_seed(0)
    images = torch.randn(32, 3, 224, 224, generator=g)
    targets = torch.randn(32, 66, generator=g) * 9.5 + 5.75
    return sd, tokens, images, targets


def _build(sd, tokens, which):
    """The oracle model (on the GPU, fp32) or the product model from the same weights and the same DoRA
    initialisation stream (A then B per layer, vision blocks first: NEW:443-445, 492-513)."""
    import hba
    from oracle import clip_ref, dora_ref
    from src.models.CLIPs.clip_hba import clip as pclip
    if which == "oracle":
        model = dora_ref.CLIPHBARef(clip_ref.build_model(sd), tokens)
        layer = dora_ref.DoRALayerRef
    else:
        model = dora_ref.CLIPHBARef(pclip.build_model(sd), tokens)
        layer = hba.DoRALayer
    torch.manual_seed(123)
    dora_ref.apply_dora_ref(model, 2, 1, r=32, layer_cls=layer)
    dora_ref.switch_dora_ref(model, layer_cls=layer)
    model = model.to(DEV)
    # the prompts are a plain attribute (NEW:282): keep them on the device so that the per-call `.to(device)` of
    # the wrapper is a no-op - a host->device copy is not allowed inside a captured step
    model.tokenized_prompts = model.tokenized_prompts.to(DEV)
    return model


def _trainable(model):
    return [(n, p) for n, p in model.named_parameters() if p.requires_grad]


@pytest.fixture(scope="module")
def oracle_step(clip_state):
    """Oracle forward / backward at ViT-L/14, batch 32 (fp32 on the GPU, TF32 off): computed once."""
    _exact_fp32()
    sd, tokens, images, targets = clip_state
    oracle = _build(sd, tokens, "oracle")
    pred = oracle(images.to(DEV))
    loss = torch.nn.MSELoss()(pred, targets.to(DEV))
    loss.backward()
    out = {"pred": pred.detach().cpu(), "loss": float(loss),
           "grads": {n: p.grad.detach().cpu() for n, p in _trainable(oracle)}}
    del oracle
    torch.cuda.empty_cache()
    return out


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_clip_hba_vitl14_batch32_matches_oracle(clip_state, oracle_step, precision):
    import hba
    sd, tokens, images, targets = clip_state
    hba.set_precision(precision)
    try:
        product = _build(sd, tokens, "product")
        pred = product(images.to(DEV))
        loss = torch.nn.MSELoss()(pred, targets.to(DEV))
        loss.backward()
        torch.cuda.synchronize()
        names = [n for n, _ in _trainable(product)]
        assert names == list(oracle_step["grads"]) and len(names) == 9
        assert sum(p.numel() for _, p in _trainable(product)) == 183040      # RUNLOG:57
        m = {"pred_rel_max": rel_max(pred.cpu(), oracle_step["pred"]),
             "loss_rel": abs(float(loss) - oracle_step["loss"]) / abs(oracle_step["loss"]),
             "grad_rel_max": {}, "grad_rel_l2": {}, "grad_cos": {}}
        for n, p in _trainable(product):
            want = oracle_step["grads"][n]
            m["grad_rel_max"][n] = rel_max(p.grad.cpu(), want)
            m["grad_rel_l2"][n] = rel_l2(p.grad.cpu(), want)
            m["grad_cos"][n] = cosine(p.grad.cpu(), want)
        _record(f"clip_hba_vitl14_b32_{precision}", m)
        if precision == "fp32":
            assert m["pred_rel_max"] < 1e-3, m
            assert m["loss_rel"] < 1e-3, m
            for n in names:
                assert m["grad_rel_max"][n] < 1e-3, (n, m["grad_rel_max"][n])
        else:
            assert m["pred_rel_max"] < BF16_TOL["pred_rel_max"], m
            assert m["loss_rel"] < BF16_TOL["loss_rel"], m
            for n in names:
                assert m["grad_rel_l2"][n] < BF16_TOL["grad_rel_l2"], (n, m["grad_rel_l2"][n])
                assert m["grad_cos"][n] > BF16_TOL["grad_cos"], (n, m["grad_cos"][n])
    finally:
        hba.set_precision("bf16")
        del product
        torch.cuda.empty_cache()


def test_clip_hba_bf16_training_trajectory_matches_fp32_oracle(clip_state):
    """24 optimisation steps (6 passes over 4 batches of 32) in bf16 mode through the product's own step
    (functions._pipeline_core.TrainStep: forward, MSE, NaN guard, backward, fused AdamW) against the oracle
    trained with torch.optim.AdamW in fp32: per-step loss and the behavioural-RSA rho of 48 held-out images."""
    import hba
    from functions import _pipeline_core as core
    from hba import rsa
    _exact_fp32()
    sd, tokens, _, _ = clip_state
    g = torch.Generator().manual_seed(7)
    images = torch.randn(128, 3, 224, 224, generator=g)
    targets = torch.randn(128, 66, generator=g) * 9.5 + 5.75
    probe = torch.randn(48, 3, 224, 224, generator=g)
    ref_rdm = 1 - np.corrcoef(np.random.default_rng(2).standard_normal((48, 66)))
    np.fill_diagonal(ref_rdm, 0)
    steps, every = 24, 8

    def rho_of(model):
        with torch.no_grad():
            emb = torch.cat([model(probe[i:i + 24].to(DEV)) for i in (0, 24)], 0)
        from scipy.stats import spearmanr
        e = emb.double().cpu().numpy()
        rdm = 1 - np.corrcoef(e)
        np.fill_diagonal(rdm, 0)
        iu = np.triu_indices(48, k=1)
        return float(spearmanr(ref_rdm[iu], rdm[iu])[0]), emb

    # oracle: torch autograd + torch.optim.AdamW over model.parameters() (NEW:1181)
    oracle = _build(sd, tokens, "oracle")
    opt = torch.optim.AdamW(oracle.parameters(), lr=3e-4)
    crit = torch.nn.MSELoss()
    want_loss, want_rho = [], []
    for s in range(steps):
        if s % every == 0:
            want_rho.append(rho_of(oracle)[0])
        b = (s % 4) * 32
        opt.zero_grad()
        loss = crit(oracle(images[b:b + 32].to(DEV)), targets[b:b + 32].to(DEV))
        loss.backward()
        opt.step()
        want_loss.append(float(loss))
    want_rho.append(rho_of(oracle)[0])
    del oracle, opt
    torch.cuda.empty_cache()

    hba.set_precision("bf16")
    product = _build(sd, tokens, "product")
    popt = core.make_optimizer(product, 3e-4)
    step = core.TrainStep(product, popt, crit, DEV)
    got_loss, got_rho, got_rho_gpu = [], [], []
    evaluator = rsa.RSAEvaluator(ref_rdm, DEV)
    for s in range(steps):
        if s % every == 0:
            r, emb = rho_of(product)
            got_rho.append(r)
            got_rho_gpu.append(evaluator(emb)[0])
        b = (s % 4) * 32
        step(images[b:b + 32].to(DEV), targets[b:b + 32].to(DEV))
        got_loss.append(float(step.last_loss))
    r, emb = rho_of(product)
    got_rho.append(r)
    got_rho_gpu.append(evaluator(emb)[0])
    assert int(step.guard.total) == 0
    loss_err = [abs(a - b) / abs(b) for a, b in zip(got_loss, want_loss)]
    rho_err = [abs(a - b) for a, b in zip(got_rho, want_rho)]
    _record("clip_hba_bf16_trajectory", {"loss_oracle": want_loss, "loss_product": got_loss,
                                         "rho_oracle": want_rho, "rho_product": got_rho,
                                         "max_loss_rel": max(loss_err), "max_rho_abs": max(rho_err)})
    assert want_loss[-1] < want_loss[3], "the oracle's loss did not move: the trajectory pins nothing"
    assert max(loss_err) < BF16_TOL["traj_loss_rel"], loss_err
    assert max(rho_err) < BF16_TOL["traj_rho_abs"], rho_err
    # the product's own RSA tail (libhba RDM / ranks / Pearson) on the product's embeddings: rho within 1e-4
    # of scipy on the same embeddings (north star), in fact ~1e-12
    assert max(abs(a - b) for a, b in zip(got_rho_gpu, got_rho)) < 1e-4
    del product
    torch.cuda.empty_cache()


# ------------------------------------------------------------------------------ ViT-B/16, batch 256
@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_vit_b16_batch256_matches_oracle(precision):
    """timm-shaped ViT-B/16 (VIT:283), batch 256 x 224^2, 1000 classes, cross entropy (VIT:139): logits, loss and
    the gradient of every parameter against oracle/vit_ref.py run in fp32 on the same GPU."""
    import hba
    from hba import vit
    from oracle import vit_ref
    _exact_fp32()
    ref = vit_ref.create_model("vit_base_patch16_224", num_classes=1000, seed=3)
    with torch.no_grad():     # biases / LayerNorm parameters away from their 0 / 1 initial values
        for p_ in ref.parameters():
            if p_.ndim == 1:
                p_.add_(torch.randn_like(p_) * 0.1)
    prod = vit.create_model("vit_base_patch16_224", num_classes=1000)
    prod.load_state_dict(ref.state_dict(), strict=True)
    ref, prod = ref.to(DEV), prod.to(DEV)
    g = torch.Generator().manual_seed(0)
    x = torch.randn(256, 3, 224, 224, generator=g).to(DEV)
    y = torch.randint(0, 1000, (256,), generator=g).to(DEV)
    want = ref(x)
    lo = torch.nn.functional.cross_entropy(want, y)
    lo.backward()
    hba.set_precision(precision)
    try:
        out = prod(x)
        lp = torch.nn.functional.cross_entropy(out, y)
        lp.backward()
        torch.cuda.synchronize()
        m = {"logits_rel_max": rel_max(out, want), "loss_rel": abs(float(lp) - float(lo)) / abs(float(lo)),
             "grad_rel_max": 0.0, "grad_rel_l2": 0.0, "grad_cos_min": 1.0}
        worst = None
        for (n, a), (_, b) in zip(prod.named_parameters(), ref.named_parameters()):
            assert a.grad is not None, n
            e = rel_max(a.grad, b.grad)
            if e > m["grad_rel_max"]:
                m["grad_rel_max"], worst = e, n
            m["grad_rel_l2"] = max(m["grad_rel_l2"], rel_l2(a.grad, b.grad))
            m["grad_cos_min"] = min(m["grad_cos_min"], cosine(a.grad, b.grad))
        m["worst_grad"] = worst
        _record(f"vit_b16_b256_{precision}", m)
        if precision == "fp32":
            assert m["logits_rel_max"] < 1e-3 and m["loss_rel"] < 1e-3, m
            assert m["grad_rel_max"] < 1e-3, m
        else:
            assert m["logits_rel_max"] < 3e-2 and m["loss_rel"] < 5e-3, m
            assert m["grad_rel_l2"] < 5e-2 and m["grad_cos_min"] > 0.995, m
    finally:
        hba.set_precision("bf16")
        del ref, prod
        torch.cuda.empty_cache()


# ------------------------------------------------------------------------------ RSA at scale (config 5)
@pytest.mark.parametrize("ties", [False, True])
def test_rsa_at_scale_1854_matches_numpy_scipy(ties):
    """N = 1,854 embeddings x 66 -> 1,717,731 pairs: RDM within 1e-12 of numpy.corrcoef (float64), ranks
    bit-exact against scipy.stats.rankdata('average'), rho within 1e-10 of scipy.stats.spearmanr.
    `ties`: embeddings quantised so that thousands of RDM entries coincide exactly (tie runs)."""
    from scipy.stats import rankdata, spearmanr
    from hba import ops, rsa
    rng = np.random.default_rng(11 + ties)
    N = 1854
    emb = rng.standard_normal((N, 66)).astype(np.float32)
    if ties:
        emb = np.round(emb[:, :66] * 2) / 2          # coarse grid: many identical rows / correlations
        emb[100:140] = emb[60:100]                   # 40 duplicated embeddings
    ref = 1 - np.corrcoef(rng.standard_normal((N, 66)))
    np.fill_diagonal(ref, 0)
    iu = np.triu_indices(N, k=1)
    want_rdm = 1 - np.corrcoef(emb)                  # float32 in -> float64 out (numpy promotes)
    np.fill_diagonal(want_rdm, 0)
    ev = rsa.RSAEvaluator(ref, DEV)
    rho, p, rdm = ev(torch.from_numpy(emb).to(DEV))
    assert float(np.abs(rdm - want_rdm).max()) < 1e-12
    # ranks of the product's own RDM vector: bit-exact against scipy on the identical doubles
    tri = torch.from_numpy(np.ascontiguousarray(rdm[iu])).to(DEV)
    ranks = torch.empty_like(tri)
    ops.rank_avg_f64(tri, ranks)
    want_ranks = rankdata(rdm[iu], method="average")
    if ties:
        assert len(np.unique(want_ranks)) < len(want_ranks) - 1000, "the tie case produced no ties"
    assert np.array_equal(ranks.cpu().numpy(), want_ranks)
    want_rho = spearmanr(ref[iu], rdm[iu])[0]
    _record(f"rsa_n1854_ties{int(ties)}", {"rho": rho, "rho_scipy": float(want_rho),
                                           "abs_diff": abs(rho - float(want_rho))})
    assert abs(rho - want_rho) < 1e-10
    assert abs(rho - spearmanr(ref[iu], want_rdm[iu])[0]) < 1e-4     # north star: rho within 1e-4
    assert 0.0 <= p <= 1.0
