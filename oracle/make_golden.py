"""ORACLE (test infrastructure only).  Generates tests/golden/*.pt by running the REFERENCE'S OWN code
(imported read-only from /root/reference through oracle/ref_loader.py, with the un-vendored ``clip``
dependency replaced by the restated oracle/clip_ref.py) on seeded synthetic inputs, on the CPU.

Run here (the GPU box has no /root/reference):   python oracle/make_golden.py
Fixtures:
  dora_layer.pt        reference DoRALayer: init (m, D), merged weight, autograd grads        NEW:407-463
  rsa_tail.pt          reference behavioral_RSA on fixed embeddings (rho, p, RDM)              NEW:605-654
  shuffle_targets.pt   reference shuffle_targets with a seeded CPU generator                   NEW:731-779
  tiny_clip_forward.pt reference CLIPHBA + apply_dora_to_ViT forward / loss / grads (ViT-tiny) NEW:268-304, 484-544
  tiny_training.pt     reference train_model, 3 epochs, `uniform_images` window on epoch 2     NEW:782-1063
"""
import csv
import os
import sys
import tempfile

import numpy as np
import scipy.io
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import clip_ref, ref_loader  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


from oracle.synth import PROMPTS, ListDataset, synthetic_problem  # noqa: E402


def build_reference_model(NEW, n_vis=2, n_txt=1, r=8, seed=123):
    model = NEW.CLIPHBA(PROMPTS, backbone_name="ViT-tiny/14", pos_embedding=True)
    torch.manual_seed(seed)
    NEW.apply_dora_to_ViT(model, n_vision_layers=n_vis, n_transformer_layers=n_txt, r=r, dora_dropout=0.1)
    NEW.switch_dora_layers(model, freeze_all=True, dora_state=True)
    return model


def main():
    os.makedirs(OUT, exist_ok=True)
    NEW, BASE = ref_loader.load_reference()

    # ---- DoRALayer
    torch.manual_seed(7)
    lin = torch.nn.Linear(64, 48)
    torch.manual_seed(8)
    layer = NEW.DoRALayer(lin, r=8)
    G = torch.randn(48, 64, generator=torch.Generator().manual_seed(9))
    W = layer.weight
    (W * G).sum().backward()
    torch.save({"lin_weight": lin.weight.detach(), "lin_bias": lin.bias.detach(), "m": layer.m.detach(),
                "D": layer.D, "A": layer.delta_D_A.detach(), "B": layer.delta_D_B.detach(),
                "scaling": layer.scaling, "W": W.detach(), "G": G, "dm": layer.m.grad,
                "dA": layer.delta_D_A.grad, "dB": layer.delta_D_B.grad}, os.path.join(OUT, "dora_layer.pt"))

    # ---- behavioral_RSA tail (identity "model": the loader yields the embeddings themselves)
    with tempfile.TemporaryDirectory() as tmp:
        g = torch.Generator().manual_seed(11)
        emb = torch.randn(48, 66, generator=g) * 3 + 1
        emb[5] = emb[4]  # identical rows -> exact ties in the RDM
        human = 1 - np.corrcoef(torch.randn(48, 20, generator=g).numpy())
        np.fill_diagonal(human, 0)
        human = np.round(human, 2)  # ties on the reference side too
        mat = os.path.join(tmp, "RDM48_triplet.mat")
        scipy.io.savemat(mat, {"RDM48_triplet": human})
        ds = ListDataset([(f"img{i}", emb[i]) for i in range(48)], RDM48_triplet_dir=mat)
        loader = torch.utils.data.DataLoader(ds, batch_size=32, shuffle=False)
        ident = torch.nn.Identity()
        rho, p, rdm = NEW.behavioral_RSA(ident, loader, torch.device("cpu"))
        torch.save({"emb": emb, "human_rdm": human, "rho": float(rho), "p": float(p), "model_rdm": rdm},
                   os.path.join(OUT, "rsa_tail.pt"))

    # ---- shuffle_targets
    t = torch.arange(40, dtype=torch.float32).reshape(8, 5)
    gen = torch.Generator().manual_seed(42 + 3 * 1000 + 2)
    torch.save({"targets": t, "seed": 42 + 3 * 1000 + 2,
                "shuffled": NEW.shuffle_targets(t, generator=gen)}, os.path.join(OUT, "shuffle_targets.pt"))

    # ---- tiny CLIP-HBA forward / backward
    prob = synthetic_problem()
    model = build_reference_model(NEW)
    crit = torch.nn.MSELoss()
    x, y = prob["train_images"][:3], prob["train_targets"][:3]
    pred = model(x)
    loss = crit(pred, y)
    loss.backward()
    torch.save({"n_images": 3, "pred": pred.detach(), "loss": float(loss),
                "grads": {n: p.grad.clone() for n, p in model.named_parameters() if p.requires_grad},
                "dora_init": {n: p.detach().clone() for n, p in model.named_parameters() if p.requires_grad},
                "n_trainable": NEW.count_trainable_parameters(model)},
               os.path.join(OUT, "tiny_clip_forward.pt"))

    # ---- 3 epochs of the reference train_model (perturbation: uniform_images on epoch 2)
    with tempfile.TemporaryDirectory() as tmp:
        mat = os.path.join(tmp, "RDM48_triplet.mat")
        scipy.io.savemat(mat, {"RDM48_triplet": prob["human_rdm"]})
        NEW.seed_everything(1)
        model = build_reference_model(NEW)
        # NEW.save_dora_parameters hard-codes ViT-L/14's block numbers (22/23/11, NEW:665-669); for the
        # 3/2-block miniature the checkpoint writer is replaced by the located-layers variant (same
        # on-disk format).  Everything numerical below is the reference's own train_model.
        NEW.save_dora_parameters = lambda m, path, epoch, logger=None: \
            ref_loader._save_dora_parameters_stub(m, path, epoch, 2, 1)
        tr = ListDataset([(f"tr{i}", prob["train_images"][i], prob["train_targets"][i]) for i in range(16)])
        te = ListDataset([(f"te{i}", prob["test_images"][i], prob["test_targets"][i]) for i in range(8)])
        rs = ListDataset([(f"rs{i}", prob["rsa_images"][i]) for i in range(8)], RDM48_triplet_dir=mat)
        gen = torch.Generator()
        gen.manual_seed(1)
        tl = torch.utils.data.DataLoader(tr, batch_size=8, shuffle=True, generator=gen)
        el = torch.utils.data.DataLoader(te, batch_size=8, shuffle=False)
        rl = torch.utils.data.DataLoader(rs, batch_size=8, shuffle=False)
        opt = torch.optim.AdamW(model.parameters(), lr=3e-4)
        logger = NEW.setup_logger(os.path.join(tmp, "log.txt"))
        res = os.path.join(tmp, "res.csv")
        NEW.train_model(model, tl, el, rl, torch.device("cpu"), opt, crit, epochs=3, training_res_path=res,
                        training_run=2, perturb_length=1, perturb_seed=42, mean=5.75, std=9.5,
                        perturb_distribution="target", perturb_type="uniform_images", logger=logger,
                        early_stopping_patience=10, dora_parameters_path=os.path.join(tmp, "dora"),
                        random_state_path=os.path.join(tmp, "rand"), dataloader_generator=gen)
        rows = list(csv.reader(open(res)))
        ck = torch.load(os.path.join(tmp, "dora", "epoch3_dora_params.pth"))
        rstate = torch.load(os.path.join(tmp, "rand", "epoch3_random_states.pth"), weights_only=False)
        torch.save({"csv_rows": rows, "dora_epoch3": ck,
                    "optimizer_state_keys": sorted(rstate["optimizer_state_dict"]["state"].keys()),
                    "random_state_keys": sorted(rstate.keys())}, os.path.join(OUT, "tiny_training.pt"))
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
