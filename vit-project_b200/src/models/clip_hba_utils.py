"""Plug-in for the un-vendored ``src.models.clip_hba_utils`` the reference's baseline pipeline imports
(BASE:22, used at BASE:683-690).  On-disk format identical to the perturbation pipeline's own
``save_dora_parameters`` (NEW:657-693): {<module path>.{m,delta_D_A,delta_D_B}: cpu tensor}."""
import os

import torch


def save_dora_parameters(model, path, epoch, vision_layers, transformer_layers, log_fn=None):
    root = model.module if isinstance(model, torch.nn.DataParallel) else model
    vb = root.clip_model.visual.transformer.resblocks
    tb = root.clip_model.transformer.resblocks
    paths = [f"clip_model.visual.transformer.resblocks.{len(vb) - vision_layers + i}.attn.out_proj"
             for i in range(vision_layers)]
    paths += [f"clip_model.transformer.resblocks.{len(tb) - transformer_layers + i}.attn.out_proj"
              for i in range(transformer_layers)]
    state = {}
    for p in paths:
        mod = root
        for attr in p.split("."):
            mod = getattr(mod, attr)
        for name in ("m", "delta_D_A", "delta_D_B"):
            state[f"{p}.{name}"] = getattr(mod, name).detach().cpu()
    os.makedirs(path, exist_ok=True)
    target = os.path.join(path, f"epoch{epoch + 1}_dora_params.pth")
    torch.save(state, target)
    if log_fn:
        log_fn(f"DoRA parameters saved: {target}")
