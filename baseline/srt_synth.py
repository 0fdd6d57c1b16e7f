"""Synthetic batches and model construction for the reference's SRT encoder/decoder (the CALLERS of the hot path):
`TransformingSRT` built from the reference's own YAML, fed random images / rays / poses with the shapes the datasets
produce (source/data/nvs/multishapenet.py:150-275, clevr_tr.py) — there is no dataset on the GPU box (TensorFlow / sunds
are absent) and the hot path does not depend on pixel content.  Used by bench.py (`--workload train_step`, BASELINE
config 5) and by tests/test_reference_modules.py.  Not imported by the product package.
"""
from __future__ import annotations

import copy
from typing import Dict

import torch

from gta_b200.synth import patch_coords, random_extrinsics

SHAPES = {
    # run: (input views, target views, image H, W, encoder token grid h, w, target points per view)
    "msn/GTA/gta_so3": (5, 5, 128, 128, 16, 16, 512),        # runs/msn/GTA/gta_so3/config.yaml (num_points 2560 // 5)
    "clevrtr/GTA/gta": (2, 3, 120, 160, 15, 20, 853),        # runs/clevrtr/GTA/gta/config.yaml (240x320 / 2, 2560 // 3)
}


def shapes_for(run: str):
    ds = run.split("/")[0]
    for k, v in SHAPES.items():
        if k.split("/")[0] == ds:
            return v
    raise KeyError(run)


def build_model(ref, run: str, device, dropout: float | None = None):
    """TransformingSRT(cfg['model']['args']) exactly as train.py:174-178 builds it."""
    cfg = copy.deepcopy(__import__("baseline.ref_loader", fromlist=["config"]).config(run))
    args = cfg["model"]["args"]
    if dropout is not None:
        args["encoder_kwargs"]["dropout"] = dropout
        args["decoder_kwargs"]["dropout"] = dropout
    model = ref.models_nvs.TransformingSRT(args).to(device)
    return model, cfg


def make_batch(run: str, B: int, device, seed: int = 0, dtype=torch.float32) -> Dict[str, torch.Tensor]:
    """Keys and shapes of one collated training batch (source/trainer.py:85-100)."""
    Ni, Nt, H, W, h, w, P = shapes_for(run)
    g = torch.Generator().manual_seed(seed)
    r = lambda *s: torch.rand(*s, generator=g)
    n = lambda *s: torch.randn(*s, generator=g)
    unit = lambda x: x / x.norm(dim=-1, keepdim=True)
    d = {
        "input_images": r(B, Ni, 3, H, W),
        "input_camera_pos": n(B, Ni, 3),
        "input_rays": unit(n(B, Ni, H, W, 3)),
        "target_pixels": r(B, Nt, P, 3),
        "target_camera_pos": n(B, Nt, P, 3),
        "target_rays": unit(n(B, Nt, P, 3)),
        "input_transforms": random_extrinsics(g, B, Ni),                       # view 0 canonical (= identity)
        "target_transforms": random_extrinsics(g, B, Nt, first_identity=False),
        "input_coord": torch.from_numpy(patch_coords(h, w))[None, None].expand(B, Ni, h * w, 2).contiguous(),
        "target_coord": r(B, Nt, P, 2),
    }
    return {k: v.to(device=device, dtype=dtype) for k, v in d.items()}


def loss_fn(model, batch, mixed_prec: bool):
    """SRTTrainer.compute_loss (source/trainer.py:85-125) on a device-resident batch."""
    extras = {k: batch[k] for k in ("input_transforms", "target_transforms", "input_coord", "target_coord")}
    extras["input_rays"], extras["target_rays"] = batch["input_rays"], batch["target_rays"]
    with torch.autocast(device_type="cuda", dtype=torch.bfloat16, enabled=mixed_prec):
        pred, extras = model(batch["input_images"], batch["input_camera_pos"], batch["input_rays"],
                             batch["target_camera_pos"], batch["target_rays"], extras)
    tgt = batch["target_pixels"].flatten(1, 2)
    pred = pred.reshape(*tgt.shape)
    return ((pred.float() - tgt) ** 2).mean((1, 2)).mean(0), pred
