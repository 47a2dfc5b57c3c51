"""Full-size checks (BASELINE.json configurations: ViT-L/14 + DoRA r32 at batch 32; ViT-B/16 at batch 256):
the CPU oracle needs minutes for these sizes, so parity is checked through properties that do not depend on
the size - batch-permutation equivariance, the cosine-logit bound, fp32-mode vs bf16-mode agreement, bit
identity of the trunk-cached / graph-replayed / host-launched forms of the same step, the known loss of a
freshly initialised classifier (ln C) and the consistency of the gradient buckets."""
import math
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda:0")


class _Args:
    batch, backbone, precision = 32, "ViT-L/14", "bf16"


@pytest.fixture(scope="module")
def clip_hba_full():
    import bench
    import hba
    hba.set_precision("bf16")
    model, opt = bench.build_gpu_model(_Args, DEV)
    g = torch.Generator().manual_seed(0)
    images = torch.randn(32, 3, 224, 224, generator=g).to(DEV)
    targets = (torch.randn(32, 66, generator=g) * 9.5 + 5.75).to(DEV)
    yield model, opt, images, targets
    hba.set_precision("bf16")


def test_full_size_forward_properties(clip_hba_full):
    import hba
    model, _, images, _ = clip_hba_full
    eng = model.clip_model.hba_engine()
    eng.trunk_cache = None
    with torch.no_grad():
        pred = model(images)
        assert pred.shape == (32, 66) and bool(torch.isfinite(pred).all())
        # logits = exp(logit_scale) * cos(image, text): |pred| <= 100 (the weights mimic the pretrained scale)
        assert float(pred.abs().max()) <= 100.0 * (1 + 1e-3)
        # every image is processed independently: permuting the batch permutes the predictions, bit for bit
        perm = torch.randperm(32, generator=torch.Generator().manual_seed(1)).to(DEV)
        assert torch.equal(model(images[perm]), pred[perm])
        # a batch of 5 (ragged last tile of every GEMM) gives the same rows
        assert torch.equal(model(images[:5]), pred[:5])
        # fp32 mode (3-pass bf16 products, fp32 attention) against bf16 mode: stated bf16 tolerance
        hba.set_precision("fp32")
        try:
            pred32 = model(images)
        finally:
            hba.set_precision("bf16")
        err = float((pred - pred32).abs().max()) / float(pred32.abs().max())
        assert err < 6e-2, err   # 24 blocks of bf16 operands against the fp32-mode path


def test_full_size_step_forms_are_bit_identical(clip_hba_full, monkeypatch):
    """One optimisation step at ViT-L/14, batch 32: host-launched, graph-replayed (text tower on a
    parallel branch) and trunk-cached forms produce the same loss trajectory and DoRA parameters."""
    from functions import _pipeline_core as core
    model, opt, images, targets = clip_hba_full
    eng = model.clip_model.hba_engine()
    crit = torch.nn.MSELoss()
    params = [p for p in model.parameters() if p.requires_grad]
    start = [p.detach().clone() for p in params]
    ids = list(range(100, 132))
    ids_dev = torch.tensor(ids, device=DEV)

    def run(graph, cached):
        for p, s0 in zip(params, start):
            p.data.copy_(s0)
        opt2 = core.make_optimizer(model, 3e-4)
        monkeypatch.setenv("HBA_STEP_GRAPH", "1" if graph else "0")
        eng.trunk_cache = None
        if cached:
            core.enable_trunk_cache(model, 256)
        step = core.TrainStep(model, opt2, crit, DEV)
        losses = []
        for _ in range(4):
            if cached:
                step(images, targets, ids, ids_dev)
            else:
                step(images, targets)
            losses.append(float(step.last_loss))
        return losses, [p.detach().clone() for p in params], int(step.guard.total)

    base = run(graph=False, cached=False)
    assert base[2] == 0 and base[0][-1] != base[0][0]          # no skipped batch, parameters moved
    for graph, cached in ((True, False), (False, True), (True, True)):
        got = run(graph, cached)
        assert got[0] == base[0], (graph, cached, got[0], base[0])
        for a, b in zip(got[1], base[1]):
            assert torch.equal(a, b)
    eng.trunk_cache = None


def test_vit_b16_full_size_step_properties():
    """ViT-B/16, batch 256, 1000 classes: a fresh classifier's loss is ln(1000); the flat gradient buffer is
    finite, fully written and consistent with the per-parameter views; the captured step equals the
    host-launched one bit for bit."""
    import hba
    from hba import vit
    hba.set_precision("bf16")
    g = torch.Generator(device=DEV).manual_seed(0)
    images = torch.randn(256, 3, 224, 224, device=DEV, generator=g)
    labels = torch.randint(0, 1000, (256,), device=DEV, generator=g)
    results = []
    for use_graph in (False, True):
        torch.manual_seed(0)
        model = vit.create_model("vit_base_patch16_224", num_classes=1000).to(DEV)
        tr = vit.DataParallelTrainer(model, lr=0.1, momentum=0.9, weight_decay=1e-4, use_graph=use_graph)
        losses = []
        for i in range(3):
            loss, hits = tr.step(images, labels)
            losses.append(float(loss))
            if i == 0:
                flat = tr.eng.flat_grad
                assert bool(torch.isfinite(flat).all()) and float(flat.abs().max()) > 0
                for p in model.parameters():          # every parameter received a gradient
                    assert float(tr.eng.grad_of[id(p)].abs().max()) > 0, "a parameter without gradient"
                assert 0 <= int(hits) <= 256
        assert abs(losses[0] - math.log(1000.0)) < 0.35   # timm init: near-uniform predictions
        results.append((losses, [p.detach().clone() for p in model.parameters()]))
        del model, tr
        torch.cuda.empty_cache()
    assert results[0][0] == results[1][0]
    for a, b in zip(results[0][1], results[1][1]):
        assert torch.equal(a, b)
