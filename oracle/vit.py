"""Oracle restatement of the DINOv2 ViT the MapAnything encoder wraps (test infrastructure only).

Follows the vendored copy in the reference (paths relative to
/root/reference/mapanything/models/external/dinov2/):
  models/vision_transformer.py:57-198  module layout / hyper-parameters
  models/vision_transformer.py:208-242 positional-embedding interpolation
  models/vision_transformer.py:244-265 token preparation (cls + pos)
  models/vision_transformer.py:292-310 forward_features -> x_norm_patchtokens
  layers/patch_embed.py:65-87, layers/block.py:93-119, layers/attention.py:53-70, layers/mlp.py:32-40,
  layers/layer_scale.py:25-26
  hub/backbones.py:21-66,91-99         dinov2_vitl14 defaults (img 518, patch 14, init_values 1.0, mlp FFN)
State-dict keys are identical to the vendored model so one synthetic checkpoint drives both.
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn
import torch.nn.functional as F


# The reference's attention modules call F.scaled_dot_product_attention (dinov2/layers/attention.py:73-90, the uniception /
# pi3-style blocks likewise).  The oracle's DEFAULT is the explicit softmax below -- plain fp32 arithmetic, what the golden
# vectors pin.  bench.py's GPU-eager baseline sets USE_SDPA = True so that the timed path is what PyTorch dispatches on the
# GPU (cuDNN / flash SDPA); both forms are the same function up to rounding.
USE_SDPA = False


class OracleAttention(nn.Module):
    def __init__(self, dim: int, num_heads: int):
        super().__init__()
        self.num_heads = num_heads
        self.qkv = nn.Linear(dim, 3 * dim, bias=True)
        self.proj = nn.Linear(dim, dim, bias=True)

    def forward(self, x, scale_factor: float = 1.0):
        b, n, c = x.shape
        hd = c // self.num_heads
        qkv = self.qkv(x).view(b, n, 3, self.num_heads, hd).permute(2, 0, 3, 1, 4)
        q, k, v = qkv[0], qkv[1], qkv[2]
        if USE_SDPA:
            o = F.scaled_dot_product_attention(q, k, v, scale=hd**-0.5 * scale_factor)
            return self.proj(o.transpose(1, 2).reshape(b, n, c))
        att = torch.softmax((q * (hd**-0.5 * scale_factor)) @ k.transpose(-1, -2), dim=-1)
        return self.proj((att @ v).transpose(1, 2).reshape(b, n, c))


class OracleMlp(nn.Module):
    def __init__(self, dim: int, hidden: int):
        super().__init__()
        self.fc1 = nn.Linear(dim, hidden)
        self.fc2 = nn.Linear(hidden, dim)

    def forward(self, x):
        return self.fc2(F.gelu(self.fc1(x)))  # exact (erf) GELU, as nn.GELU()


class OracleLayerScale(nn.Module):
    def __init__(self, dim: int):
        super().__init__()
        self.gamma = nn.Parameter(torch.ones(dim))

    def forward(self, x):
        return x * self.gamma


class OracleBlock(nn.Module):
    """Pre-LN block: x += ls1(attn(norm1(x))); x += ls2(mlp(norm2(x)))."""

    def __init__(self, dim: int, num_heads: int, mlp_ratio: float = 4.0, layer_scale: bool = True):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim, eps=1e-6)
        self.attn = OracleAttention(dim, num_heads)
        self.ls1 = OracleLayerScale(dim) if layer_scale else nn.Identity()
        self.norm2 = nn.LayerNorm(dim, eps=1e-6)
        self.mlp = OracleMlp(dim, int(dim * mlp_ratio))
        self.ls2 = OracleLayerScale(dim) if layer_scale else nn.Identity()

    def forward(self, x, scale_factor: float = 1.0):
        x = x + self.ls1(self.attn(self.norm1(x), scale_factor))
        return x + self.ls2(self.mlp(self.norm2(x)))


class OraclePatchEmbed(nn.Module):
    def __init__(self, patch: int, in_chans: int, dim: int):
        super().__init__()
        self.proj = nn.Conv2d(in_chans, dim, kernel_size=patch, stride=patch)

    def forward(self, x):
        return self.proj(x).flatten(2).transpose(1, 2)  # (B, h*w, C), row-major over the patch grid


class OracleDinoV2(nn.Module):
    def __init__(self, img_size=518, patch_size=14, embed_dim=1024, depth=24, num_heads=16, mlp_ratio=4.0,
                 interpolate_offset=0.1):
        super().__init__()
        self.patch_size = patch_size
        self.embed_dim = embed_dim
        self.interpolate_offset = interpolate_offset
        n = (img_size // patch_size) ** 2
        self.cls_token = nn.Parameter(torch.zeros(1, 1, embed_dim))
        self.pos_embed = nn.Parameter(torch.zeros(1, n + 1, embed_dim))
        self.mask_token = nn.Parameter(torch.zeros(1, embed_dim))  # unused at inference; kept for key parity
        self.patch_embed = OraclePatchEmbed(patch_size, 3, embed_dim)
        self.blocks = nn.ModuleList([OracleBlock(embed_dim, num_heads, mlp_ratio) for _ in range(depth)])
        self.norm = nn.LayerNorm(embed_dim, eps=1e-6)

    def pos_embed_for(self, n_patches: int, w: int, h: int) -> torch.Tensor:
        """vision_transformer.py:208-242 (argument order (w, h) = (x.shape[2], x.shape[3]) as in the reference)."""
        n0 = self.pos_embed.shape[1] - 1
        if n_patches == n0 and w == h:
            return self.pos_embed
        pe = self.pos_embed.float()
        cls_pe, patch_pe = pe[:, :1], pe[:, 1:]
        dim = pe.shape[-1]
        w0, h0 = w // self.patch_size, h // self.patch_size
        m = int(math.sqrt(n0))
        assert m * m == n0
        if self.interpolate_offset:
            kw = {"scale_factor": (float(w0 + self.interpolate_offset) / m, float(h0 + self.interpolate_offset) / m)}
        else:
            kw = {"size": (w0, h0)}
        grid = F.interpolate(patch_pe.reshape(1, m, m, dim).permute(0, 3, 1, 2), mode="bicubic", antialias=False, **kw)
        assert grid.shape[-2:] == (w0, h0)
        return torch.cat([cls_pe, grid.permute(0, 2, 3, 1).reshape(1, -1, dim)], dim=1)

    def tokens(self, img: torch.Tensor) -> torch.Tensor:
        _, _, w, h = img.shape
        x = self.patch_embed(img)
        x = torch.cat([self.cls_token.expand(x.shape[0], -1, -1), x], dim=1)
        return x + self.pos_embed_for(x.shape[1] - 1, w, h).to(x.dtype)

    def forward_patch_tokens(self, img: torch.Tensor) -> torch.Tensor:
        """== forward_features(img)["x_norm_patchtokens"]: (B, N, C)."""
        x = self.tokens(img)
        for blk in self.blocks:
            x = blk(x)
        return self.norm(x)[:, 1:]


class OracleDinoV2Encoder(nn.Module):
    """uniception DINOv2Encoder: hub model under `.model`, output (B, C, H/14, W/14)."""

    def __init__(self, name="dinov2_large", data_norm_type="dinov2", size="large", with_registers=False,
                 gradient_checkpointing=False, torch_hub_force_reload=False, vit_kwargs=None, **_):
        super().__init__()
        assert size == "large" and not with_registers, "only the ViT-L/14 no-register encoder is on the hot path"
        self.name = name
        self.data_norm_type = data_norm_type
        self.model = OracleDinoV2(**(vit_kwargs or {}))
        self.patch_size = self.model.patch_size
        self.enc_embed_dim = self.model.embed_dim

    def forward(self, image: torch.Tensor, data_norm_type: str) -> torch.Tensor:
        assert data_norm_type == self.data_norm_type, "image normalisation does not match the encoder"
        b, _, h, w = image.shape
        assert h % self.patch_size == 0 and w % self.patch_size == 0
        tok = self.model.forward_patch_tokens(image)
        return tok.permute(0, 2, 1).reshape(b, -1, h // self.patch_size, w // self.patch_size).contiguous()
