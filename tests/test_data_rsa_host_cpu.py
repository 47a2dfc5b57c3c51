"""CPU tests of host-side pieces without a device dependency: the Spearman p-value (hba.rsa), and the HBM-resident
input store (hba.data) - device-agnostic host logic: the resident
loader visits the items exactly like the reference's `DataLoader(dataset, batch_size, shuffle, generator=...)`
(NEW:1119-1126) and consumes the shuffle generator identically, so that the generator state saved / restored per
epoch (NEW:129-131, 709-727) stays interchangeable."""
import pytest
import torch
from torch.utils.data import DataLoader, Dataset, Subset


class _Things(Dataset):
    """(name, image, target) items like ThingsDataset (NEW:196-204)."""

    def __init__(self, n):
        g = torch.Generator().manual_seed(n)
        self.images = torch.randn(n, 3, 4, 4, generator=g)
        self.targets = torch.randn(n, 66, generator=g)

    def __len__(self):
        return len(self.images)

    def __getitem__(self, i):
        return f"img_{i:03d}.jpg", self.images[i], self.targets[i]


@pytest.mark.parametrize("shuffle", [True, False])
def test_resident_loader_equals_dataloader_order_and_generator_consumption(shuffle):
    from hba import data
    ds = _Things(23)
    store = data.ResidentStore(ds, "cpu")
    assert len(store) == 23 and store.images.dtype == torch.float32 and store.names[5] == "img_005.jpg"
    g_ref, g_mine = torch.Generator().manual_seed(11), torch.Generator().manual_seed(11)
    ref = DataLoader(ds, batch_size=4, shuffle=shuffle, generator=g_ref)
    mine = data.ResidentLoader(store, 4, shuffle=shuffle, generator=g_mine)
    assert len(mine) == len(ref) == 6
    for epoch in range(3):
        got, want = list(mine), list(ref)
        assert len(got) == len(want)
        for (gn, gi, gt), (wn, wi, wt) in zip(got, want):
            assert list(gn) == list(wn) and torch.equal(gi, wi) and torch.equal(gt, wt)
        assert torch.equal(g_mine.get_state(), g_ref.get_state())          # same consumption, epoch after epoch
        if epoch == 0:       # restoring a saved generator state (NEW:129-131) replays the same epoch on both sides
            saved = g_ref.get_state().clone()
    g_ref.set_state(saved)
    g_mine.set_state(saved)
    assert [list(b[0]) for b in mine] == [list(b[0]) for b in ref]


def test_resident_loader_subset_ids_and_inference_items():
    from hba import data
    ds = _Things(12)
    store = data.ResidentStore(ds, "cpu")
    subset = [7, 2, 9, 4, 0]
    g_ref, g_mine = torch.Generator().manual_seed(3), torch.Generator().manual_seed(3)
    ref = DataLoader(Subset(ds, subset), batch_size=2, shuffle=True, generator=g_ref)   # SubsetWithIndices, NEW:164-177
    mine = data.ResidentLoader(store, 2, shuffle=True, generator=g_mine, index_map=subset)
    it = iter(mine)
    for wn, wi, wt in ref:
        gn, gi, gt = next(it)
        assert list(gn) == list(wn) and torch.equal(gi, wi) and torch.equal(gt, wt)
        assert mine.last_ids == [data.image_id(n, store.namespace) for n in gn] == mine.last_ids_dev.tolist()
    assert list(it) == [] and mine.last_ids is None
    assert data.image_id("img_007.jpg") == data.image_id("img_007.jpg") != data.image_id("img_008.jpg")
    # the same file name under another image directory is another image (another cache entry)
    assert data.image_id("img_007.jpg", "/data/a") != data.image_id("img_007.jpg", "/data/b")
    other = data.ResidentStore(_Things(12), "cpu")
    assert other.namespace != store.namespace and not set(other.ids) & set(store.ids)
    # (name, image) items - ThingsInferenceDataset, NEW:206-224 - give two-element batches
    pairs = [(f"v{i}.jpg", torch.full((3, 2, 2), float(i))) for i in range(5)]
    inf = data.ResidentLoader(data.ResidentStore(pairs, "cpu"), 4)
    batches = list(inf)
    assert [len(b) for b in batches] == [2, 2] and list(batches[1][0]) == ["v4.jpg"] and batches[0][1].shape == (4, 3, 2, 2)
    assert list(data.ResidentLoader(data.ResidentStore(ds, "cpu"), 4, index_map=[])) == []


@pytest.mark.parametrize("n", [3, 10, 66, 1128, 1_717_731])
def test_spearman_pvalue_equals_scipy(n):
    """hba.rsa.spearman_pvalue (the one scalar of the RSA tail evaluated on the host, NEW:652 / MEAS:351) against the
    p-value `scipy.stats.spearmanr` attaches to the same rho, incl. rho = +-1 and tiny samples."""
    import numpy as np
    from scipy import stats
    from hba.rsa import spearman_pvalue
    rng = np.random.default_rng(n)
    m = min(n, 5000)                                    # scipy on a sample of the same size class; rho is what matters
    for _ in range(5):
        a = rng.standard_normal(m)
        b = a * rng.uniform(-1, 1) + rng.standard_normal(m) * rng.uniform(0.01, 2)
        rho, p = stats.spearmanr(a, b)
        assert spearman_pvalue(float(rho), m) == pytest.approx(float(p), rel=1e-10, abs=1e-300)
    assert spearman_pvalue(1.0, n) == 0.0 and spearman_pvalue(-1.0, n) == 0.0      # t = +-inf
    rho1, p1 = stats.spearmanr(np.arange(m), np.arange(m))          # scipy's rho of identical rankings: 1 - 1 ulp
    if m > 3:
        assert spearman_pvalue(float(rho1), m) == pytest.approx(float(p1), rel=1e-6, abs=1e-300)
    assert spearman_pvalue(0.0, n) == pytest.approx(1.0)
    assert np.isnan(spearman_pvalue(0.5, 2))            # dof = 0: scipy returns nan as well


@pytest.mark.parametrize("shape,r", [((96, 80), 16), ((64, 64), 8), ((128, 256), 32)])
def test_product_dora_layer_construction_equals_oracle_and_reference(shape, r):
    """hba.DoRALayer.__init__ runs on the host (the merge is the device part): decomposition S = ||W0^T||_col,
    D = W0^T / S, Kaiming-uniform A then B on the global RNG, frozen bias, attribute names and state_dict keys
    (NEW:407-445) - bit for bit against the oracle restatement and, where mounted, the reference's own class."""
    import hba
    from oracle import dora_ref, ref_loader
    in_f, out_f = shape
    torch.manual_seed(3)
    lin = torch.nn.Linear(in_f, out_f)
    torch.manual_seed(4)
    mine = hba.DoRALayer(lin, r=r)
    rng_after = torch.get_rng_state()
    torch.manual_seed(4)
    want = dora_ref.DoRALayerRef(lin, r=r)
    assert torch.equal(torch.get_rng_state(), rng_after)                      # same consumption of the global RNG
    for n in ("m", "D", "delta_D_A", "delta_D_B", "bias"):
        assert torch.equal(getattr(mine, n), getattr(want, n)), n
    assert mine.scaling == want.scaling and mine.original_layer is lin
    assert sorted(k for k in mine.state_dict() if not k.startswith("original_layer")) == \
        sorted(k for k in want.state_dict() if not k.startswith("original_layer"))
    assert [n for n, p in mine.named_parameters() if p.requires_grad and not n.startswith("original_layer")] == \
        [n for n, p in want.named_parameters() if p.requires_grad and not n.startswith("original_layer")]
    assert (mine.in_features, mine.out_features) == (in_f, out_f) if hasattr(mine, "in_features") else True
    if ref_loader.reference_available():
        NEW, BASE = ref_loader.load_reference()
        for cls in (NEW.DoRALayer, BASE.DoRALayer):
            torch.manual_seed(4)
            ref = cls(lin, r=r)
            for n in ("m", "D", "delta_D_A", "delta_D_B", "bias"):
                assert torch.equal(getattr(mine, n), getattr(ref, n)), n
            assert sorted(mine.state_dict()) == sorted(ref.state_dict())
            assert {n: p.requires_grad for n, p in mine.named_parameters()} == {n: p.requires_grad for n, p in ref.named_parameters()}


def test_trunk_cache_bookkeeping_on_the_host():
    """hba.engine.TrunkCache (frozen-trunk activation cache, north star item 2) - the host-side bookkeeping with a
    stand-in engine: miss -> store -> hit returns the stored rows in request order, partial presence is a miss,
    a changed weight stamp / precision / device drops every entry, ids outside the capacity are refused."""
    import types
    from hba.engine import TrunkCache
    bufs = {}

    def _buf(name, shape, dtype=torch.float32):
        return bufs.setdefault((name, tuple(shape)), torch.empty(*shape, dtype=dtype))

    eng = types.SimpleNamespace(vis=types.SimpleNamespace(T=5, d=4), precision="bf16", _stamp=1,
                                device=torch.device("cpu"), _buf=_buf)
    cache = TrunkCache(capacity=16)
    ctx = cache.lookup(eng, [3, 7])
    assert ctx["hit"] is False and "x" not in ctx and not cache.all_present([3])
    g = torch.Generator().manual_seed(0)
    x, a = torch.randn(2 * 5, 4, generator=g), torch.randn(2 * 5, 4, generator=g)
    cache.store(ctx["ids"], x, a)
    assert cache.all_present([7, 3]) and not cache.all_present([3, 8])
    hit = cache.lookup(eng, [7, 3])                                   # request order, not insertion order
    assert hit["hit"] and torch.equal(hit["x"], torch.cat([x[5:], x[:5]])) and torch.equal(hit["a"], torch.cat([a[5:], a[:5]]))
    assert hit["x"].shape == (10, 4)
    dev = cache.lookup_device_ids(eng, torch.tensor([3, 3, 7]))
    assert dev["hit"] and dev["ids"] is None and torch.equal(dev["x"][:5], x[:5]) and torch.equal(dev["x"][10:], x[5:])
    assert cache.lookup(eng, [3, 8])["hit"] is False                  # one absent image: the whole batch recomputes
    with pytest.raises(RuntimeError, match="capacity"):
        cache.lookup(eng, [16])
    with pytest.raises(RuntimeError, match="capacity"):
        cache.lookup(eng, [-1])
    eng._stamp = 2                                                    # weights restaged (e.g. another checkpoint loaded)
    assert cache.lookup(eng, [3, 7])["hit"] is False and not cache.all_present([3])
    cache.store(cache.lookup(eng, [3, 7])["ids"], x, a)
    eng.precision = "fp32"                                            # precision mode changed
    assert cache.lookup(eng, [3, 7])["hit"] is False


# ------------------------------------------------------------------------------- round 2 host pieces
def test_find_dora_checkpoints_and_reference_rdm(tmp_path):
    """hba.rsa_scale (RSA at scale over sweep checkpoints, BASELINE config 5): the checkpoint walk follows the
    directory layouts of SWEEP:198-207 / LEN:128-137, sorted by run directory and epoch NUMBER (epoch10 after
    epoch9), and the full-set reference RDM is 1 - corrcoef of the behavioural embedding with a zero diagonal."""
    import numpy as np
    from hba import rsa_scale
    root = tmp_path
    for rel in ("base/dora/epoch2_dora_params.pth", "base/dora/epoch10_dora_params.pth",
                "base/dora/epoch9_dora_params.pth", "out/random_target_e1_l2/dora_params_1/epoch3_dora_params.pth",
                "out/training_run7/dora_params_run7/epoch8_dora_params.pth", "out/training_run7/other.pth",
                "base/rand/epoch2_random_states.pth"):
        p = root / rel
        p.parent.mkdir(parents=True, exist_ok=True)
        p.write_bytes(b"x")
    files = [str(p)[len(str(root)) + 1:] for p in map(str, rsa_scale.find_dora_checkpoints(str(root)))]
    assert files == ["base/dora/epoch2_dora_params.pth", "base/dora/epoch9_dora_params.pth",
                     "base/dora/epoch10_dora_params.pth",
                     "out/random_target_e1_l2/dora_params_1/epoch3_dora_params.pth",
                     "out/training_run7/dora_params_run7/epoch8_dora_params.pth"]
    t = np.random.default_rng(0).standard_normal((7, 66))
    rdm = rsa_scale.reference_rdm_from_targets(t)
    assert rdm.shape == (7, 7) and np.allclose(np.diag(rdm), 0) and np.allclose(rdm, rdm.T)
    assert np.allclose(rdm[0, 1], 1 - np.corrcoef(t[0], t[1])[0, 1])


def test_auto_k_slices_for_row_gemms_and_weight_gradients():
    """Split-K plans (hba.ops.auto_k_slices): a weight-gradient GEMM keeps >= 8 k-blocks per slice, a row GEMM
    (CLS / EOT rows: one row tile) may go down to 4, and a problem that already fills the 74 CTA pairs is not split."""
    from hba import ops
    assert ops.auto_k_slices(1024, 1024, 8224) > 1            # dW of a 1024 x 1024 out_proj: 16 tiles
    assert ops.auto_k_slices(9472, 512, 1024) == 1            # 74 tiles = one full round: nothing to gain
    for (M, N, K) in [(32, 1024, 4096), (66, 768, 3072), (32, 768, 1024)]:
        s8, s4 = ops.auto_k_slices(M, N, K), ops.auto_k_slices(M, N, K, min_kblocks=4)
        assert 1 <= s8 <= s4 and s4 > 1
        assert (K // 64) // s4 >= 4


def test_trunk_cache_grows_when_a_later_run_brings_more_images(monkeypatch):
    """The frozen CLIP (and its engine) is reused by every run of a process: a second run on another image directory
    draws ids beyond the first cache's capacity.  `enable_trunk_cache` then installs a larger, empty cache and drops
    the graphs captured on the model (they gather out of the old buffers) instead of failing the first lookup."""
    import types
    from functions import _pipeline_core as core
    from hba import data
    eng = types.SimpleNamespace(trunk_cache=None)
    model = torch.nn.Module()
    model.clip_model = types.SimpleNamespace(hba_engine=lambda: eng)
    monkeypatch.setattr(data, "_NAME_IDS", {("a", f"img{i}"): i for i in range(100)})
    core.enable_trunk_cache(model, 100)
    first = eng.trunk_cache
    assert first.capacity == 164
    model.__dict__["_hba_train_step"] = object()
    model.__dict__["_hba_forward_graphs"] = object()
    core.enable_trunk_cache(model, 100)                          # same image set again: nothing changes
    assert eng.trunk_cache is first and "_hba_train_step" in model.__dict__
    data._NAME_IDS.update({("b", f"img{i}"): 100 + i for i in range(100)})
    core.enable_trunk_cache(model, 100)                          # another directory: ids 100..199 > capacity 164
    assert eng.trunk_cache is not first and eng.trunk_cache.capacity == 264 and not eng.trunk_cache.present
    assert "_hba_train_step" not in model.__dict__ and "_hba_forward_graphs" not in model.__dict__
    core.enable_trunk_cache(torch.nn.Module(), 10)               # a model without a libhba engine: no-op
