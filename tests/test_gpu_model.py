"""End-to-end parity of the libhba CLIP-HBA forward / backward against the CPU oracle
(oracle/clip_ref.py + oracle/dora_ref.py) on identical seeded weights and inputs.

Tolerances (north star): fp32 mode — predictions and gradients within 1e-3 relative;
bf16 mode — stated here as 3e-2 of max|pred| for predictions and 8e-2 relative for gradients."""
import copy

import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"


def rel_err(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def build_pair(name="ViT-tiny/14", n_vis=2, n_txt=1, r=8, n_prompts=5, product_layer="hba"):
    import hba
    from oracle import clip_ref, dora_ref
    from src.models.CLIPs.clip_hba import clip as pclip
    sd = clip_ref.synthetic_state_dict(name, seed=1)
    prompts = ["metallic; artificial", "food-related", "animal-related", "textile", "plant-related",
               "house-related; furnishing-related"][:n_prompts]
    tokens = torch.stack([clip_ref.tokenize(p) for p in prompts])  # [S,1,77] like NEW:282
    oracle = dora_ref.CLIPHBARef(clip_ref.build_model(sd), tokens)
    torch.manual_seed(123)
    dora_ref.apply_dora_ref(oracle, n_vis, n_txt, r=r)
    dora_ref.switch_dora_ref(oracle)
    product = dora_ref.CLIPHBARef(pclip.build_model(sd), tokens)
    layer_cls = hba.DoRALayer if product_layer == "hba" else dora_ref.DoRALayerRef
    torch.manual_seed(123)
    dora_ref.apply_dora_ref(product, n_vis, n_txt, r=r, layer_cls=layer_cls)
    dora_ref.switch_dora_ref(product, layer_cls=layer_cls)
    product.to(DEV)
    return oracle, product


def dora_params(model):
    return [(n, p) for n, p in model.named_parameters() if p.requires_grad]


@pytest.mark.parametrize("precision,tol_pred,tol_grad", [("fp32", 1e-3, 1e-3), ("bf16", 3e-2, 8e-2)])
@pytest.mark.parametrize("product_layer", ["hba", "reference_style"])
def test_tiny_forward_backward_parity(precision, tol_pred, tol_grad, product_layer):
    import hba
    hba.set_precision(precision)
    oracle, product = build_pair(product_layer=product_layer)
    g = torch.Generator().manual_seed(0)
    images = torch.randn(3, 3, 224, 224, generator=g)
    targets = torch.randn(3, 5, generator=g) * 9.5 + 5.75
    crit = torch.nn.MSELoss()
    po = oracle(images)
    lo = crit(po, targets)
    lo.backward()
    pp = product(images.to(DEV))
    lp = crit(pp, targets.to(DEV))
    lp.backward()
    assert pp.shape == po.shape == (3, 5)
    e = rel_err(pp, po.detach())
    assert e < tol_pred, ("pred", e)
    assert abs(float(lp) - float(lo)) / abs(float(lo)) < tol_pred * 3
    names_o, names_p = dora_params(oracle), dora_params(product)
    assert [n for n, _ in names_o] == [n for n, _ in names_p] and len(names_o) == 9
    for (n, a), (_, b) in zip(names_p, names_o):
        assert a.grad is not None, n
        e = rel_err(a.grad, b.grad)
        assert e < tol_grad, (n, e)
    hba.set_precision("bf16")


def test_no_grad_eval_and_batch_sizes():
    import hba
    hba.set_precision("fp32")
    oracle, product = build_pair()
    g = torch.Generator().manual_seed(5)
    for B in (1, 4):
        images = torch.randn(B, 3, 224, 224, generator=g)
        with torch.no_grad():
            po = oracle(images)
            pp = product(images.to(DEV))
        assert rel_err(pp, po) < 1e-3
    hba.set_precision("bf16")


def test_training_steps_track_oracle():
    """5 AdamW steps: loss trajectory and final DoRA parameters match (fp32 mode)."""
    import hba
    from hba.optim import FusedAdamW
    hba.set_precision("fp32")
    oracle, product = build_pair()
    opt_o = torch.optim.AdamW(oracle.parameters(), lr=3e-4)
    opt_p = FusedAdamW(product.parameters(), lr=3e-4)
    crit = torch.nn.MSELoss()
    g = torch.Generator().manual_seed(1)
    for step in range(5):
        images = torch.randn(2, 3, 224, 224, generator=g)
        targets = torch.randn(2, 5, generator=g) * 9.5 + 5.75
        opt_o.zero_grad()
        lo = crit(oracle(images), targets)
        lo.backward()
        opt_o.step()
        opt_p.zero_grad()
        lp = crit(product(images.to(DEV)), targets.to(DEV))
        lp.backward()
        opt_p.step()
        assert abs(float(lp) - float(lo)) / abs(float(lo)) < 1e-3, (step, float(lp), float(lo))
    for (n, a), (_, b) in zip(dora_params(product), dora_params(oracle)):
        assert rel_err(a.detach(), b.detach()) < 1e-3, n
    # state_dict layout interchangeable with torch.optim.AdamW (resume fidelity, NEW:124-126)
    so, sp = opt_o.state_dict(), opt_p.state_dict()
    assert so["param_groups"][0]["params"] == sp["param_groups"][0]["params"]
    assert set(so["state"].keys()) == set(sp["state"].keys())
    hba.set_precision("bf16")
