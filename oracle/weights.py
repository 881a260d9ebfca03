"""Deterministic synthetic checkpoints (test infrastructure only).

There is no network for real checkpoints, so parity runs use random weights.  To let the reference code
(imported from /root/reference when fixtures are generated), the oracle and the CUDA path all load THE SAME
values independent of module construction order, every tensor is drawn from its own generator seeded by
crc32(key) ^ seed, with a scale chosen by the key's role so activations stay O(1) through ~50 layers:
  *norm*.weight  1 + 0.1 n     *norm*.bias / *.bias  0.02 n     *.gamma (LayerScale)  1 + 0.1 n
  cls_token / pos_embed / scale_token  0.02 n     mask_token 0
  weights with >= 2 dims  n / sqrt(fan_in)       (fan_in = prod(shape[1:]); ConvTranspose: shape[0] * k*k / stride^2)
"""
from __future__ import annotations

import zlib
from typing import Dict

import torch


def _randn(key: str, shape, seed: int) -> torch.Tensor:
    g = torch.Generator(device="cpu").manual_seed((zlib.crc32(key.encode()) ^ (seed * 0x9E3779B1)) & 0x7FFFFFFF)
    return torch.randn(tuple(shape), generator=g, dtype=torch.float32)


def synth_tensor(key: str, shape, seed: int = 0, transposed_conv: bool = False) -> torch.Tensor:
    leaf = key.split(".")[-1]
    parent = key.split(".")[-2] if "." in key else ""
    is_norm = "norm" in parent
    if leaf == "mask_token":
        return torch.zeros(tuple(shape))
    if leaf in ("cls_token", "pos_embed", "scale_token"):
        return 0.02 * _randn(key, shape, seed)
    if leaf == "gamma" or (is_norm and leaf == "weight"):
        return 1.0 + 0.1 * _randn(key, shape, seed)
    if leaf == "bias":
        return 0.02 * _randn(key, shape, seed)
    if len(shape) >= 2:
        if transposed_conv:  # (Cin, Cout, k, k) with k == stride: each output pixel sees Cin inputs
            fan_in = shape[0]
        else:
            fan_in = 1
            for s in shape[1:]:
                fan_in *= s
        return _randn(key, shape, seed) / fan_in**0.5
    return 0.02 * _randn(key, shape, seed)


def synth_state_dict(model: torch.nn.Module, seed: int = 0) -> Dict[str, torch.Tensor]:
    """One synthetic value per state-dict key of `model` (aliased keys such as dense_head.0.* get the value of
    their canonical name because nn.Module.state_dict() lists the shared tensor under each alias and
    load_state_dict copies key by key; we therefore draw by CANONICAL key)."""
    tconv = {
        name + ".weight" for name, mod in model.named_modules() if isinstance(mod, torch.nn.ConvTranspose2d)
    }
    sd = {}
    for key, val in model.state_dict().items():
        canon = canonical_key(key)
        sd[key] = synth_tensor(canon, val.shape, seed, transposed_conv=(key in tconv or canon in tconv)).to(val.dtype)
    return sd


def canonical_key(key: str) -> str:
    if key.startswith("dense_head.0."):
        return "dpt_feature_head." + key[len("dense_head.0."):]
    if key.startswith("dense_head.1."):
        return "dpt_regressor_head." + key[len("dense_head.1."):]
    return key


def load_synthetic(model: torch.nn.Module, seed: int = 0) -> torch.nn.Module:
    model.load_state_dict(synth_state_dict(model, seed), strict=True)
    return model


@torch.no_grad()
def init_reference_style(model: torch.nn.Module, seed: int = 0) -> torch.nn.Module:
    """"Random-init weights of that architecture" as the reference would have them without a checkpoint:
    DINOv2's own init (vision_transformer.py:203-207, :384-389: trunc-normal(0.02) Linear weights, zero biases,
    pos_embed trunc-normal(0.02), cls_token N(0, 1e-6), LayerScale gamma = init_values = 1.0 per hub/backbones.py:25),
    the same timm-style init for the multi-view transformer, scale_token trunc-normal(0.02) (model.py:201-202), and
    PyTorch's default reset_parameters() for every conv / linear of the heads and geometric encoders."""
    torch.manual_seed(seed)
    nn = torch.nn
    for name, mod in model.named_modules():
        vit_like = name.startswith("encoder.model") or name.startswith("info_sharing")
        if isinstance(mod, nn.Linear):
            if vit_like:
                nn.init.trunc_normal_(mod.weight, std=0.02)
                if mod.bias is not None:
                    nn.init.zeros_(mod.bias)
            else:
                mod.reset_parameters()
        elif isinstance(mod, (nn.Conv2d, nn.ConvTranspose2d)):
            mod.reset_parameters()
        elif isinstance(mod, nn.LayerNorm):
            mod.reset_parameters()
    for name, p in model.named_parameters():
        leaf = name.split(".")[-1]
        if leaf == "gamma":
            p.fill_(1.0)
        elif leaf in ("pos_embed", "scale_token"):
            nn.init.trunc_normal_(p, std=0.02)
        elif leaf == "cls_token":
            nn.init.normal_(p, std=1e-6)
        elif leaf == "mask_token":
            p.zero_()
    return model
