"""GPU parity tests of the ViT-B/16 data-parallel baseline path (reference VIT = Training/vit_training/
baseline/train_vit_sgd.py) against the CPU oracle oracle/vit_ref.py, through the C-ABI:
MN-major GEMM operands (dX / dW without transposes), ragged K, column sums, LayerNorm parameter
gradients, the full attention backward, and the whole forward / backward / SGD step."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu

DEV = "cuda"


def _imports():
    import hba
    from hba import ops
    return hba, ops


def rel_err(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def make_operand(ops, x, split):
    rows, cols = x.shape
    op = ops.Operand.empty(rows, cols, split, x.device, zero=True)
    ops.split_bf16(x.contiguous(), op)
    return op


def operand_value(op, rows=None, cols=None):
    rows, cols = rows or op.rows, cols or op.K
    v = op.buf[:rows, :cols].double()
    if op.lo_off > 0:
        v = v + op.buf[:rows, op.lo_off:op.lo_off + cols].double()
    return v


@pytest.mark.parametrize("M,N,K", [(128, 256, 64), (256, 256, 128), (384, 128, 200), (768, 3072, 591),
                                   (1000, 768, 32), (72, 1000, 3)])
@pytest.mark.parametrize("a_mn,b_mn", [(True, True), (False, True), (True, False)])
@pytest.mark.parametrize("split", [False, True])
def test_gemm_mn_major_and_ragged_k(M, N, K, a_mn, b_mn, split):
    """C = A B^T with A and/or B stored [K, M] / [K, N] (MN-major UMMA descriptors) and K not a
    multiple of 64 (TMA zero fill): dW = dY^T X and dX = dY W of the ViT backward."""
    hba, ops = _imports()
    g = torch.Generator().manual_seed(M + 3 * N + 7 * K)
    pad8 = lambda n: (n + 7) // 8 * 8
    a = torch.randn(M, K, generator=g).to(DEV)
    b = torch.randn(N, K, generator=g).to(DEV)
    if a_mn:   # stored [K, pad8(M)]
        at = torch.zeros(K, pad8(M), device=DEV)
        at[:, :M] = a.t()
        A = make_operand(ops, at, split)
        A = ops.Operand(A.buf, K, M, A.lo_off)
        a_val = operand_value(A, K, M).t()
    else:      # K-major operand: ragged K needs zero padding to a multiple of 64 in fp32 mode
        Kp = (K + 63) // 64 * 64
        ap = torch.zeros(M, Kp, device=DEV)
        ap[:, :K] = a
        A = make_operand(ops, ap, split)
        a_val = operand_value(A)[:, :K]
    if b_mn:
        bt = torch.zeros(K, pad8(N), device=DEV)
        bt[:, :N] = b.t()
        Bo = make_operand(ops, bt, split)
        Bo = ops.Operand(Bo.buf, K, N, Bo.lo_off)
        b_val = operand_value(Bo, K, N).t()
    else:
        Kp = (K + 63) // 64 * 64
        bp = torch.zeros(N, Kp, device=DEV)
        bp[:, :K] = b
        Bo = make_operand(ops, bp, split)
        b_val = operand_value(Bo)[:, :K]
    out = torch.full((M, pad8(N)), float("nan"), device=DEV)[:, :N]
    ops.gemm(A, Bo, M if not a_mn else None, a_mn=a_mn, b_mn=b_mn, K=K, out_f32=out)
    torch.cuda.synchronize()
    if split:
        want, tol = a.double() @ b.double().t(), 2e-5
    else:
        want, tol = a_val @ b_val.t(), 2e-6 * math.sqrt(max(K, 64))
    assert torch.isfinite(out).all()
    assert rel_err(out, want) < tol, (rel_err(out, want), tol)


@pytest.mark.parametrize("M,N,K,slices", [(768, 768, 5000, 7), (768, 3072, 2048, 2), (256, 128, 1000, 16),
                                          (1000, 768, 777, "auto"), (768, 768, 50432, "auto")])
@pytest.mark.parametrize("split", [False, True])
def test_gemm_split_k(M, N, K, slices, split):
    """dW = dY^T X with the K range split into slices (deterministic fixed-order reduction): equal to
    the exact product of the bf16 operands and bit-identical from run to run."""
    hba, ops = _imports()
    g = torch.Generator().manual_seed(M + N + K)
    pad8 = lambda n: (n + 7) // 8 * 8
    at = torch.zeros(K, pad8(M), device=DEV)
    at[:, :M] = torch.randn(K, M, generator=g).to(DEV)
    bt = torch.zeros(K, pad8(N), device=DEV)
    bt[:, :N] = torch.randn(K, N, generator=g).to(DEV)
    A = make_operand(ops, at, split)
    A = ops.Operand(A.buf, K, M, A.lo_off)
    Bo = make_operand(ops, bt, split)
    Bo = ops.Operand(Bo.buf, K, N, Bo.lo_off)
    ws = torch.empty(16 * M * N, device=DEV)
    outs = []
    for _ in range(2):
        out = torch.full((M, N), float("nan"), device=DEV)
        ops.gemm(A, Bo, a_mn=True, b_mn=True, K=K, out_f32=out, k_slices=slices, k_workspace=ws)
        outs.append(out)
    ref = torch.empty(M, N, device=DEV)
    ops.gemm(A, Bo, a_mn=True, b_mn=True, K=K, out_f32=ref)
    torch.cuda.synchronize()
    assert torch.equal(outs[0], outs[1])
    if split:
        want, tol = at[:, :M].double().t() @ bt[:, :N].double(), 4e-5   # bf16x3 drops lo.lo (~2^-16 per product)
    else:
        want, tol = operand_value(A, K, M).t() @ operand_value(Bo, K, N), 2e-6 * math.sqrt(K)
    assert rel_err(outs[0], want) < tol, (rel_err(outs[0], want), tol)
    assert rel_err(outs[0], ref) < 2e-6 * math.sqrt(K)   # same products, different fp32 summation order


@pytest.mark.parametrize("M,N,K", [(333, 512, 128), (1000, 3072, 256), (70, 256, 64)])
def test_gemm_fused_bias_gradient_column_sums(M, N, K):
    """colsum_partial: per 32-row group column sums of the bf16 output, emitted by the GELU'-epilogue
    GEMM that produces dY (the fc1 bias gradient of the ViT backward); summed with hba_colsum."""
    hba, ops = _imports()
    from hba._lib import HBA_ACT_GELU_ERF_GRAD
    g = torch.Generator().manual_seed(M + N)
    a = torch.randn(M, K, generator=g).to(DEV)
    b = (torch.randn(N, K, generator=g) * 0.2).to(DEV)
    aux = torch.randn(M, N, generator=g).to(DEV).to(torch.bfloat16)
    A, B = make_operand(ops, a, False), make_operand(ops, b, False)
    out = ops.Operand.empty(M, N, False, DEV)
    groups = (M + 31) // 32
    part = torch.full((groups, N), float("nan"), device=DEV)
    ops.gemm(A, B, M, act=HBA_ACT_GELU_ERF_GRAD, aux=aux, out=out, colsum_partial=part)
    torch.cuda.synchronize()
    got_out = out.buf[:, :N].float()
    assert torch.isfinite(part).all()
    want_part = torch.stack([got_out[32 * i:32 * i + 32].double().sum(0) for i in range(groups)])
    assert float((part.double() - want_part).abs().max()) < 1e-4 * float(got_out.abs().max()) * 32
    total = torch.empty(N, device=DEV)
    ops.colsum(part, total, torch.empty(128 * N + 64, device=DEV))
    assert rel_err(total, got_out.double().sum(0)) < 1e-5


@pytest.mark.parametrize("rows,cols", [(1, 128), (777, 768), (5000, 1000), (50432, 768)])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_colsum(rows, cols, dtype):
    hba, ops = _imports()
    g = torch.Generator().manual_seed(rows + cols)
    x = torch.randn(rows, cols, generator=g).to(DEV).to(dtype)
    out = torch.full((cols,), 3.0, device=DEV)
    ws = torch.empty(128 * cols + 64, device=DEV)
    ops.colsum(x, out, ws)
    want = x.double().sum(0)
    assert float((out.double() - want).abs().max()) < 1e-5 * float(x.double().abs().sum(0).max())
    ops.colsum(x, out, ws, accumulate=True)
    assert float((out.double() - 2 * want).abs().max()) < 2e-5 * float(x.double().abs().sum(0).max())


@pytest.mark.parametrize("rows,cols,row_step", [(300, 768, 1), (8, 128, 197), (4000, 1024, 1)])
def test_layernorm_param_grad(rows, cols, row_step):
    hba, ops = _imports()
    g = torch.Generator().manual_seed(rows)
    eps = 1e-6
    x_all = torch.randn(rows * row_step, cols, generator=g).to(DEV) * 2 + 0.5
    dy = torch.randn(rows, cols, generator=g).to(DEV)
    x = x_all[::row_step]
    xh = (x.double() - x.double().mean(1, keepdim=True)) / torch.sqrt(x.double().var(1, unbiased=False, keepdim=True) + eps)
    want = torch.cat([(dy.double() * xh).sum(0), dy.double().sum(0)])
    out = torch.empty(2 * cols, device=DEV)
    ws = torch.empty(2 * rows + 256 * cols + 64, device=DEV)
    ops.layernorm_param_grad(dy, x_all, rows, cols, eps, out, ws, row_step=row_step)
    scale = float((dy.double().abs() * xh.abs()).sum(0).max())
    assert float((out.double() - want).abs().max()) < 1e-5 * scale


@pytest.mark.parametrize("B,T,H,causal", [(2, 197, 2, False), (3, 50, 1, False), (2, 77, 3, True), (1, 257, 2, False)])
@pytest.mark.parametrize("mode", ["fp32", "bf16"])
def test_attention_bwd_full(B, T, H, causal, mode):
    """dQ, dK, dV of softmax attention against autograd on the fp64 formula."""
    if mode == "fp32" and T > 200:
        pytest.skip("fp32 K/V/dK/dV tiles of T > 200 tokens exceed 227 KB of shared memory (ViT-B/16 has T = 197)")
    hba, ops = _imports()
    g = torch.Generator().manual_seed(B * 100 + T)
    d = H * 64
    qkv = (torch.randn(B * T, 3 * d, generator=g) * 0.7)
    do = torch.randn(B * T, d, generator=g)
    if mode == "bf16":
        qkv, do = qkv.to(torch.bfloat16).float(), do.to(torch.bfloat16).float()
    ref = qkv.double().requires_grad_(True)
    q, k, v = ref.view(B, T, 3, H, 64).permute(2, 0, 3, 1, 4)
    s = (q @ k.transpose(-1, -2)) / 8.0
    if causal:
        s = s + torch.full((T, T), float("-inf"), dtype=torch.float64).triu_(1)
    o = (s.softmax(-1) @ v).permute(0, 2, 1, 3).reshape(B * T, d)
    o.backward(do.double())
    want = ref.grad
    if mode == "fp32":
        dq = torch.empty(B * T, 3 * d, device=DEV)
        ops.attention_bwd(qkv.to(DEV), B, T, H, do.to(DEV), dq, causal=causal)
        tol = 2e-4
    else:
        dq = torch.empty(B * T, 3 * d, device=DEV, dtype=torch.bfloat16)
        ops.attention_bwd(qkv.to(DEV).to(torch.bfloat16), B, T, H, do.to(DEV).to(torch.bfloat16), dq,
                          causal=causal)
        tol = 1.5e-2
    torch.cuda.synchronize()
    assert rel_err(dq.float(), want) < tol, rel_err(dq.float(), want)


@pytest.mark.parametrize("rows,cols,accumulate,lo", [(1000, 768, True, False), (37, 128, False, True),
                                                     (5000, 1024, True, False), (9, 256, True, True)])
def test_layernorm_bwd_fused(rows, cols, accumulate, lo):
    """dx, its bf16 hi/lo copy, dgamma|dbeta and colsum(dx) of the one-pass kernel vs the fp64 formulas."""
    hba, ops = _imports()
    g = torch.Generator().manual_seed(rows + cols)
    eps = 1e-6
    x = (torch.randn(rows, cols, generator=g) * 2 + 0.5).to(DEV)
    dy = torch.randn(rows, cols, generator=g).to(DEV)
    gamma = (torch.rand(cols, generator=g) + 0.5).to(DEV)
    dx0 = torch.randn(rows, cols, generator=g).to(DEV)
    xd = x.double().requires_grad_(True)
    gd = gamma.double().requires_grad_(True)
    bd = torch.zeros(cols, dtype=torch.float64, device=DEV, requires_grad=True)
    torch.nn.functional.layer_norm(xd, (cols,), gd, bd, eps).backward(dy.double())
    want_dx = xd.grad + (dx0.double() if accumulate else 0)
    dx = dx0.clone()
    op = ops.Operand.empty(rows, cols, lo, DEV)
    dgb = torch.empty(2 * cols, device=DEV)
    cs = torch.empty(cols, device=DEV)
    ws = torch.empty(2 * 148 * 3 * cols, device=DEV)
    ops.layernorm_bwd_fused(dy, x, rows, cols, gamma, eps, dx, dgb, ws, accumulate=accumulate, dx_op=op,
                            dx_colsum=cs)
    torch.cuda.synchronize()
    assert rel_err(dx, want_dx) < 2e-5
    assert rel_err(dgb[:cols], gd.grad) < 2e-5 and rel_err(dgb[cols:], bd.grad) < 2e-5
    scale = float(want_dx.abs().sum(0).max())
    assert float((cs.double() - want_dx.sum(0)).abs().max()) < 1e-5 * scale
    got = op.buf[:, :cols].float() + (op.buf[:, cols:2 * cols].float() if lo else 0)
    assert rel_err(got, want_dx) < (2e-5 if lo else 5e-3)


@pytest.mark.parametrize("B,T,H,causal", [(2, 197, 2, False), (3, 50, 1, False), (2, 77, 3, True), (1, 256, 2, False),
                                          (4, 128, 1, False), (2, 16, 1, False), (2, 130, 1, True),
                                          (27, 197, 12, False)])
def test_attention_tensor_core_fwd_lse_bwd(B, T, H, causal):
    """tcgen05 forward (+ lse) / backward pair of the bf16 training step against autograd on the fp64
    formula; (27, 197, 12) gives every CTA of the persistent kernels more than one unit."""
    hba, ops = _imports()
    g = torch.Generator().manual_seed(B * 100 + T)
    d = H * 64
    qkv = (torch.randn(B * T, 3 * d, generator=g) * 0.7).to(torch.bfloat16)
    do = torch.randn(B * T, d, generator=g).to(torch.bfloat16)
    ref = qkv.double().requires_grad_(True)
    q, k, v = ref.view(B, T, 3, H, 64).permute(2, 0, 3, 1, 4)
    s = (q @ k.transpose(-1, -2)) / 8.0
    if causal:
        s = s + torch.full((T, T), float("-inf"), dtype=torch.float64).triu_(1)
    o = (s.softmax(-1) @ v).permute(0, 2, 1, 3).reshape(B * T, d)
    o.backward(do.double())
    want_lse = torch.logsumexp(s, -1) * 1.4426950408889634   # [B, H, T], log2 domain
    out = torch.empty(B * T, d, device=DEV, dtype=torch.bfloat16)
    lse = torch.empty(B * H * T, device=DEV)
    dq = torch.full((B * T, 3 * d), float("nan"), device=DEV, dtype=torch.bfloat16)
    ops.attention_fwd_lse(qkv.to(DEV), B, T, H, out, lse, causal=causal)
    ops.attention_bwd_lse(qkv.to(DEV), B, T, H, out, do.to(DEV), lse, dq, causal=causal)
    torch.cuda.synchronize()
    assert rel_err(out.float(), o.detach()) < 1e-2
    assert float((lse.cpu().double().view(B, H, T) - want_lse.detach()).abs().max()) < 2e-3
    got, want = dq.float().cpu().double(), ref.grad
    assert bool(torch.isfinite(got).all())
    for name, sl in (("dq", slice(0, d)), ("dk", slice(d, 2 * d)), ("dv", slice(2 * d, 3 * d))):
        e = rel_err(got[:, sl], want[:, sl])
        assert e < 1.5e-2, (name, e)


def _pair(seed=3, num_classes=10):
    from hba import vit
    from oracle import vit_ref
    ref = vit_ref.create_model("vit_tiny_test", num_classes=num_classes, seed=seed)
    with torch.no_grad():
        for p in ref.parameters():
            if p.ndim == 1:
                p.add_(torch.randn_like(p) * 0.1)
    prod = vit.create_model("vit_tiny_test", num_classes=num_classes)
    prod.load_state_dict(ref.state_dict(), strict=True)   # timm parameter names on both sides
    return ref, prod.to(DEV)


@pytest.mark.parametrize("mode,tol_out,tol_grad", [("fp32", 1e-3, 1e-3), ("bf16", 3e-2, 8e-2)])
def test_vit_forward_backward_matches_oracle(mode, tol_out, tol_grad):
    hba, ops = _imports()
    hba.set_precision(mode)
    try:
        ref, prod = _pair()
        g = torch.Generator().manual_seed(0)
        x = torch.randn(3, 3, 224, 224, generator=g)
        y = torch.tensor([1, 7, 4])
        lo = torch.nn.functional.cross_entropy(ref(x), y)
        lo.backward()
        out = prod(x.to(DEV))
        lp = torch.nn.functional.cross_entropy(out, y.to(DEV))
        lp.backward()
        torch.cuda.synchronize()
        assert rel_err(out, ref(x).detach()) < tol_out
        assert abs(float(lp) - float(lo)) < tol_out * abs(float(lo))
        for (n, a), (_, b) in zip(prod.named_parameters(), ref.named_parameters()):
            assert a.grad is not None, n
            assert rel_err(a.grad, b.grad) < tol_grad, (n, rel_err(a.grad, b.grad))
    finally:
        hba.set_precision("bf16")


def test_vit_trainer_steps_match_oracle_sgd():
    """DataParallelTrainer (world 1): fused CE + backward + fused SGD over 3 steps against
    torch.optim.SGD on the oracle (VIT:132-152 in fp32): loss trajectory and final parameters."""
    hba, ops = _imports()
    from hba import vit
    from oracle import vit_ref
    hba.set_precision("fp32")
    try:
        ref, prod = _pair(seed=5)
        g = torch.Generator().manual_seed(1)
        batches = [(torch.randn(4, 3, 224, 224, generator=g), torch.randint(0, 10, (4,), generator=g))
                   for _ in range(3)]
        want = vit_ref.train_steps(ref, batches, lr=0.1, momentum=0.9, weight_decay=1e-4)
        tr = vit.DataParallelTrainer(prod, lr=0.1, momentum=0.9, weight_decay=1e-4)
        got = []
        for images, labels in batches:
            loss, hits = tr.step(images.to(DEV), labels.to(DEV))
            got.append(float(loss))
        assert got == pytest.approx(want, rel=1e-3)
        for (n, a), (_, b) in zip(prod.named_parameters(), ref.named_parameters()):
            assert rel_err(a.detach(), b.detach()) < 2e-3, (n, rel_err(a.detach(), b.detach()))
    finally:
        hba.set_precision("bf16")


@pytest.mark.parametrize("num_classes", [10, 12])
def test_vit_trainer_cuda_graph_replay_is_bit_identical_to_eager_steps(num_classes):
    """The captured step (one graph per batch shape and learning rate) replays exactly the eager step's
    kernels: 5 steps incl. a learning-rate change give bit-identical losses and parameters.  12 classes:
    every tensor size is a multiple of 4, i.e. the vectorised SGD that also refreshes the bf16 operands
    (no staging pass in the graph); 10 classes: the generic SGD + per-step staging."""
    hba, ops = _imports()
    from hba import vit
    g = torch.Generator().manual_seed(11)
    batches = [(torch.randn(4, 3, 224, 224, generator=g).to(DEV), torch.randint(0, 10, (4,), generator=g).to(DEV))
               for _ in range(5)]
    results = []
    for use_graph in (False, True):
        torch.manual_seed(3)
        model = vit.create_model("vit_tiny_test", num_classes=num_classes).to(DEV)
        tr = vit.DataParallelTrainer(model, lr=0.1, momentum=0.9, weight_decay=1e-4, use_graph=use_graph)
        losses = []
        for i, (images, labels) in enumerate(batches):
            if i == 3:
                tr.param_groups[0]["lr"] = 0.05
            loss, hits = tr.step(images, labels)
            losses.append(float(loss))
        results.append((losses, [p.detach().clone() for p in model.parameters()]))
    assert results[0][0] == results[1][0]
    for a, b in zip(results[0][1], results[1][1]):
        assert torch.equal(a, b)


def test_operand_refreshing_sgd_equals_sgd_plus_staging(monkeypatch):
    """hba_sgd_staged (vectorised, bf16 operands refreshed in the same pass) against hba_sgd_multi followed by
    the per-step staging pass: bit-identical losses and parameters over 4 steps."""
    hba, ops = _imports()
    from hba import vit
    g = torch.Generator().manual_seed(5)
    batches = [(torch.randn(4, 3, 224, 224, generator=g).to(DEV), torch.randint(0, 12, (4,), generator=g).to(DEV))
               for _ in range(4)]
    results = []
    for fused in (True, False):
        torch.manual_seed(3)
        model = vit.create_model("vit_tiny_test", num_classes=12).to(DEV)
        tr = vit.DataParallelTrainer(model, lr=0.1, momentum=0.9, weight_decay=1e-4)
        losses = []
        for images, labels in batches:
            loss, _ = tr.step(images, labels)
            losses.append(float(loss))
            if not fused:   # drop to the generic kernel from the second step on
                tr._staged = None
        assert (tr._staged is not None) == fused
        results.append((losses, [p.detach().clone() for p in model.parameters()]))
    assert results[0][0] == results[1][0]
    for a, b in zip(results[0][1], results[1][1]):
        assert torch.equal(a, b)
