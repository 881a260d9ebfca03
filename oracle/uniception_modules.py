"""Oracle restatement of the `uniception` modules MapAnything instantiates (test infrastructure only).

PARITY UNPINNED: `uniception` (pyproject.toml:26 of the reference, no version pin) is not vendored in
/root/reference, not installed and not downloadable in this environment, and the reference ships no tests
or golden vectors for it.  These modules therefore follow
  * the call contracts visible in /root/reference/mapanything/models/mapanything/model.py
    (:157-193 encoders, :299-301 info sharing, :374-388 heads, :478-483/:588 adaptors, :637-645, :812-825,
    :972-1008, :1038-1129, :1302-1338, :1449-1469, :1532-1542),
  * the hyper-parameters in /root/reference/configs/model/{encoder/dinov2_large, info_sharing/aat_ifr_24_layers,
    pred_head/dpt_pose_scale, pred_head/adaptor_config/raydirs_depth_pose_confidence_mask_scale, task/default}.yaml,
  * the vendored structural analogs /root/reference/mapanything/models/external/vggt/{models/aggregator.py:229-360,
    layers/attention.py:46-76, heads/dpt_head.py:427-567},
  * and the published UniCeption design as recorded in SURVEY.md Appendix A (A.2 - A.7), including the
    state-dict sub-key names of A.7 so that a real checkpoint maps onto these modules.
Every assumption that a real checkpoint could falsify is a constructor argument.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F

from .vit import OracleBlock


# ------------------------------------------------------------------------------------------------
# A.2 geometric-input encoders
# ------------------------------------------------------------------------------------------------
class ResidualBlock(nn.Module):
    """conv3x3 -> GELU -> conv3x3, + shortcut (identity or conv1x1), -> GELU."""

    def __init__(self, cin: int, cout: int):
        super().__init__()
        self.conv1 = nn.Conv2d(cin, cout, 3, padding=1)
        self.conv2 = nn.Conv2d(cout, cout, 3, padding=1)
        self.shortcut = nn.Identity() if cin == cout else nn.Conv2d(cin, cout, 1)

    def forward(self, x):
        y = self.conv2(F.gelu(self.conv1(x)))
        return F.gelu(y + self.shortcut(x))


class DenseRepresentationEncoder(nn.Module):
    """`dense_rep_encoder` (ray directions 3ch / depth 1ch): PixelUnshuffle(patch) -> conv3x3 -> 2 residual blocks ->
    conv1x1 -> token LayerNorm; output (B, C, H/p, W/p)."""

    def __init__(self, name: str, in_chans: int, enc_embed_dim: int, patch_size: int, apply_pe: bool = False,
                 intermediate_dims: Sequence[int] = (588, 768, 1024), encoder_str: str = "dense_rep_encoder", **_):
        super().__init__()
        assert not apply_pe, "apply_pe is false for every MapAnything task config (configs/model/task/default.yaml)"
        self.name, self.patch_size, self.enc_embed_dim = name, patch_size, enc_embed_dim
        d0, d1, d2 = intermediate_dims
        self.unshuffle = nn.PixelUnshuffle(patch_size)
        self.conv_in = nn.Conv2d(in_chans * patch_size * patch_size, d0, 3, padding=1)
        self.encoder = nn.Sequential(ResidualBlock(d0, d1), ResidualBlock(d1, d2), nn.Conv2d(d2, enc_embed_dim, 1))
        self.norm_layer = nn.LayerNorm(enc_embed_dim, eps=1e-6)

    def forward(self, data: torch.Tensor) -> torch.Tensor:
        x = self.encoder(self.conv_in(self.unshuffle(data)))
        b, c, h, w = x.shape
        tok = self.norm_layer(x.flatten(2).transpose(1, 2))
        return tok.transpose(1, 2).reshape(b, c, h, w).contiguous()


class GlobalRepresentationEncoder(nn.Module):
    """`global_rep_encoder` (quats 4 / trans 3 / log-scale 1): MLP in -> 128 -> 256 -> 512 -> C with GELU, LayerNorm."""

    def __init__(self, name: str, in_chans: int, enc_embed_dim: int, intermediate_dims: Sequence[int] = (128, 256, 512),
                 encoder_str: str = "global_rep_encoder", **_):
        super().__init__()
        self.name, self.enc_embed_dim = name, enc_embed_dim
        dims = [in_chans, *intermediate_dims, enc_embed_dim]
        layers: List[nn.Module] = []
        for i in range(len(dims) - 1):
            layers.append(nn.Linear(dims[i], dims[i + 1]))
            if i < len(dims) - 2:
                layers.append(nn.GELU())
        self.encoder = nn.Sequential(*layers)
        self.norm_layer = nn.LayerNorm(enc_embed_dim, eps=1e-6)

    def forward(self, data: torch.Tensor) -> torch.Tensor:
        return self.norm_layer(self.encoder(data))


# ------------------------------------------------------------------------------------------------
# A.3 alternating-attention multi-view transformer with intermediate feature return
# ------------------------------------------------------------------------------------------------
def sinusoid_table(n_rows: int, dim: int, base: float = 10000.0) -> torch.Tensor:
    pos = torch.arange(n_rows, dtype=torch.float64)[:, None]
    j = torch.arange(dim, dtype=torch.float64)[None, :]
    angle = pos / torch.pow(torch.tensor(base, dtype=torch.float64), 2 * torch.div(j, 2, rounding_mode="floor") / dim)
    table = torch.where((torch.arange(dim) % 2 == 0)[None, :], torch.sin(angle), torch.cos(angle))
    return table.float()


class MultiViewAlternatingAttentionTransformerIFR(nn.Module):
    """Even blocks attend globally over all V*N + T tokens, odd blocks per view over N tokens (extra tokens bypass
    them).  Returns the final normed features plus normed snapshots after the blocks in `indices`.

    Switches for what SURVEY App. A.3 marks VERIFY (a real checkpoint / the uniception source decides; the CUDA path
    has the same switches, tests/test_variants_gpu.py):
      global_attention_first   True: even blocks global (default) / False: even blocks frame-wise
      attention_pattern        "alternating" (model_type alternating_attention) / "global" (model_type global_attention,
                               configs/model/info_sharing/gat_ifr_24_layers.yaml: every block attends over all views)
      view_pe_variant          "ref_only": view 0 gets table row 0, the others nothing (default, variant (i));
                               "ref_vs_rest": view 0 row 0, every other view row 1 (variant (ii));
                               "per_view_index": view v > 0 gets row idx[v] -- random in [1, max_num_views_for_pe) when
                               use_rand_idx_pe_for_non_reference_views (aat_ifr_24_layers_w_view_pe.yaml), else v
      use_entropy_scaling      softmax scale = head_dim^-0.5 * log(n_keys) / log(entropy_scaling_ref_len)
                               (aat_ifr_*_escaling.yaml; the constant is an assumption)"""

    def __init__(self, name: str, input_embed_dim: int, indices: Sequence[int] = (11, 17), norm_intermediate: bool = True,
                 size: Optional[str] = None, depth: int = 24, dim: int = 768, num_heads: int = 12, mlp_ratio: float = 4.0,
                 distinguish_ref_and_non_ref_views: bool = True, use_pe_for_non_reference_views: bool = False,
                 use_rand_idx_pe_for_non_reference_views: bool = False, max_num_views_for_pe: int = 1000,
                 max_num_views: Optional[int] = None, gradient_checkpointing: bool = False,
                 custom_positional_encoding=None, global_attention_first: bool = True, attention_pattern: str = "alternating",
                 view_pe_variant: Optional[str] = None, use_entropy_scaling: bool = False,
                 entropy_scaling_ref_len: Optional[int] = None, **_):
        super().__init__()
        assert custom_positional_encoding is None
        self.name, self.dim, self.depth, self.num_heads = name, dim, depth, num_heads
        self.indices = list(indices)
        self.norm_intermediate = norm_intermediate
        self.distinguish_ref_and_non_ref_views = distinguish_ref_and_non_ref_views
        self.global_attention_first, self.attention_pattern = global_attention_first, attention_pattern
        if view_pe_variant is None:
            view_pe_variant = "per_view_index" if (use_rand_idx_pe_for_non_reference_views or use_pe_for_non_reference_views) \
                else "ref_only"
        self.view_pe_variant = view_pe_variant
        self.use_rand_idx = use_rand_idx_pe_for_non_reference_views
        self.max_views = int(max_num_views or max_num_views_for_pe)
        self.fixed_view_pe_indices = None  # tests: the indices of views 1..V-1 instead of random draws
        self.use_entropy_scaling = use_entropy_scaling
        self.entropy_scaling_ref_len = entropy_scaling_ref_len
        self.proj_embed = nn.Linear(input_embed_dim, dim) if input_embed_dim != dim else nn.Identity()
        self.self_attention_blocks = nn.ModuleList(
            [OracleBlock(dim, num_heads, mlp_ratio, layer_scale=False) for _ in range(depth)]
        )
        self.norm = nn.LayerNorm(dim, eps=1e-6)
        # fixed (non-learned, non-persistent) view positional table
        self.register_buffer("view_pos_table", sinusoid_table(self.max_views, dim), persistent=False)

    def view_pe_rows(self, v: int, first_view: int = 0) -> Optional[List[int]]:
        """Table row per view (None entry = no PE), for views first_view .. first_view + v - 1 of the scene."""
        if not self.distinguish_ref_and_non_ref_views and self.view_pe_variant == "ref_only":
            return None
        rows = []
        for g in range(first_view, first_view + v):
            if g == 0:
                rows.append(0)
            elif self.view_pe_variant == "ref_only":
                rows.append(None)
            elif self.view_pe_variant == "ref_vs_rest":
                rows.append(1)
            else:
                if self.fixed_view_pe_indices is not None:
                    rows.append(int(self.fixed_view_pe_indices[g - 1]))
                elif self.use_rand_idx:
                    rows.append(int(torch.randint(1, self.max_views, (1,))))
                else:
                    rows.append(g)
        return rows

    def is_global(self, i: int) -> bool:
        return self.attention_pattern == "global" or ((i % 2 == 0) == self.global_attention_first)

    def softmax_scale_factor(self, n_keys: int) -> float:
        if not self.use_entropy_scaling:
            return 1.0
        import math

        return math.log(n_keys) / math.log(self.entropy_scaling_ref_len or n_keys)

    def forward(self, features: List[torch.Tensor], additional_input_tokens: Optional[torch.Tensor] = None):
        """features: V x (B, C, h, w); additional_input_tokens: (B, C, T).
        Returns (final_feats V x (B, D, h, w), final_extra (B, D, T), [ (feats list, extra) per index ])."""
        v = len(features)
        b, c, h, w = features[0].shape
        n = h * w
        x = torch.cat([f.flatten(2).transpose(1, 2) for f in features], dim=1)  # (B, V*N, C), view-major
        t = 0
        if additional_input_tokens is not None:
            t = additional_input_tokens.shape[2]
            x = torch.cat([x, additional_input_tokens.transpose(1, 2)], dim=1)
        x = self.proj_embed(x)
        rows = self.view_pe_rows(v)
        if rows is not None:
            parts = []
            for i, r in enumerate(rows):
                xi = x[:, i * n:(i + 1) * n]
                parts.append(xi if r is None else xi + self.view_pos_table[r].to(x.dtype).view(1, 1, -1))
            x = torch.cat(parts + [x[:, v * n:]], dim=1)

        def snapshot(y):
            y = self.norm(y)
            feats = [y[:, i * n : (i + 1) * n].transpose(1, 2).reshape(b, self.dim, h, w).contiguous() for i in range(v)]
            extra = y[:, v * n :].transpose(1, 2).contiguous() if t else None
            return feats, extra

        inter = []
        for i, blk in enumerate(self.self_attention_blocks):
            if self.is_global(i):
                x = blk(x, scale_factor=self.softmax_scale_factor(x.shape[1]))
            else:
                frames = blk(x[:, : v * n].reshape(b * v, n, self.dim), scale_factor=self.softmax_scale_factor(n))
                frames = frames.reshape(b, v * n, self.dim)
                x = torch.cat([frames, x[:, v * n :]], dim=1) if t else frames
            if i in self.indices:
                if self.norm_intermediate:
                    inter.append(snapshot(x))
                else:
                    feats = [x[:, k * n : (k + 1) * n].transpose(1, 2).reshape(b, self.dim, h, w) for k in range(v)]
                    inter.append((feats, x[:, v * n :].transpose(1, 2) if t else None))
        final_feats, final_extra = snapshot(x)
        return final_feats, final_extra, inter


# ------------------------------------------------------------------------------------------------
# A.4 DPT feature head + regression processor
# ------------------------------------------------------------------------------------------------
class ResidualConvUnit(nn.Module):
    def __init__(self, ch: int):
        super().__init__()
        self.conv1 = nn.Conv2d(ch, ch, 3, padding=1)
        self.conv2 = nn.Conv2d(ch, ch, 3, padding=1)

    def forward(self, x):
        return self.conv2(F.relu(self.conv1(F.relu(x)))) + x


class FeatureFusionBlock(nn.Module):
    def __init__(self, ch: int):
        super().__init__()
        self.resConfUnit1 = ResidualConvUnit(ch)
        self.resConfUnit2 = ResidualConvUnit(ch)
        self.out_conv = nn.Conv2d(ch, ch, 1)

    def forward(self, x0, x1=None):
        out = x0 if x1 is None else x0 + self.resConfUnit1(x1)
        out = self.resConfUnit2(out)
        out = F.interpolate(out, scale_factor=2, mode="bilinear", align_corners=True)
        return self.out_conv(out)


class _Scratch(nn.Module):
    def __init__(self, layer_dims, feature_dim):
        super().__init__()
        self.layer_rn = nn.ModuleList([nn.Conv2d(d, feature_dim, 3, padding=1, bias=False) for d in layer_dims])
        self.refinenet1 = FeatureFusionBlock(feature_dim)
        self.refinenet2 = FeatureFusionBlock(feature_dim)
        self.refinenet3 = FeatureFusionBlock(feature_dim)
        self.refinenet4 = FeatureFusionBlock(feature_dim)


class DPTFeature(nn.Module):
    def __init__(self, patch_size: int, input_feature_dims: Sequence[int], feature_dim: int = 256,
                 hooks: Sequence[int] = (0, 1, 2, 3), layer_dims: Sequence[int] = (96, 192, 384, 768), **_):
        super().__init__()
        assert patch_size in (14, 16)
        self.hooks = list(hooks)
        d, ld = list(input_feature_dims), list(layer_dims)
        self.act_postprocess = nn.ModuleList(
            [
                nn.Sequential(nn.Conv2d(d[0], ld[0], 1), nn.ConvTranspose2d(ld[0], ld[0], 4, stride=4)),
                nn.Sequential(nn.Conv2d(d[1], ld[1], 1), nn.ConvTranspose2d(ld[1], ld[1], 2, stride=2)),
                nn.Sequential(nn.Conv2d(d[2], ld[2], 1)),
                nn.Sequential(nn.Conv2d(d[3], ld[3], 1), nn.Conv2d(ld[3], ld[3], 3, stride=2, padding=1)),
            ]
        )
        self.scratch = _Scratch(ld, feature_dim)

    def forward(self, list_features: List[torch.Tensor]) -> torch.Tensor:
        layers = [self.act_postprocess[i](list_features[hk]) for i, hk in enumerate(self.hooks)]
        layers = [self.scratch.layer_rn[i](l) for i, l in enumerate(layers)]
        p4 = self.scratch.refinenet4(layers[3])[:, :, : layers[2].shape[2], : layers[2].shape[3]]
        p3 = self.scratch.refinenet3(p4, layers[2])
        p2 = self.scratch.refinenet2(p3, layers[1])
        return self.scratch.refinenet1(p2, layers[0])


class DPTRegressionProcessor(nn.Module):
    def __init__(self, input_feature_dim: int, output_dim: int, hidden_dims: Optional[Sequence[int]] = None, **_):
        super().__init__()
        hidden_dims = list(hidden_dims) if hidden_dims is not None else [input_feature_dim // 2, input_feature_dim // 2]
        self.conv1 = nn.Conv2d(input_feature_dim, hidden_dims[0], 3, padding=1)
        self.conv2 = nn.Sequential(
            nn.Conv2d(hidden_dims[0], hidden_dims[1], 3, padding=1), nn.ReLU(), nn.Conv2d(hidden_dims[1], output_dim, 1)
        )

    def forward(self, x: torch.Tensor, target_output_shape: Tuple[int, int]) -> torch.Tensor:
        x = self.conv1(x)
        x = F.interpolate(x, size=tuple(target_output_shape), mode="bilinear", align_corners=True)
        return self.conv2(x)


# ------------------------------------------------------------------------------------------------
# A.6 pose head, scale head
# ------------------------------------------------------------------------------------------------
class ResConvBlock(nn.Module):
    """final_relu_after_skip=True: relu(skip + conv3(y)) (default); False: skip + relu(conv3(y)) -- the Reloc3r / ACE form
    SURVEY App. A.6 cites (a VERIFY item: only the uniception source decides; the weights are the same either way)."""

    def __init__(self, cin: int, cout: int, final_relu_after_skip: bool = True):
        super().__init__()
        self.final_relu_after_skip = final_relu_after_skip
        self.head_skip = nn.Identity() if cin == cout else nn.Conv2d(cin, cout, 1)
        self.res_conv1 = nn.Conv2d(cin, cout, 1)
        self.res_conv2 = nn.Conv2d(cout, cout, 3, padding=1)
        self.res_conv3 = nn.Conv2d(cout, cout, 1)

    def forward(self, x):
        y = F.relu(self.res_conv1(x))
        y = F.relu(self.res_conv2(y))
        y = self.res_conv3(y)
        return F.relu(self.head_skip(x) + y) if self.final_relu_after_skip else self.head_skip(x) + F.relu(y)


class PoseHead(nn.Module):
    def __init__(self, patch_size: int, input_feature_dim: int, num_resconv_block: int = 2,
                 rot_representation_dim: int = 4, final_relu_after_skip: bool = True, **_):
        super().__init__()
        c = input_feature_dim
        self.res_conv = nn.ModuleList([ResConvBlock(c, c, final_relu_after_skip) for _ in range(num_resconv_block)])
        self.more_mlps = nn.Sequential(nn.Linear(c, c), nn.ReLU(), nn.Linear(c, c), nn.ReLU())
        self.fc_t = nn.Linear(c, 3)
        self.fc_rot = nn.Linear(c, rot_representation_dim)

    def forward(self, feat: torch.Tensor) -> torch.Tensor:
        for blk in self.res_conv:
            feat = blk(feat)
        g = self.more_mlps(feat.mean(dim=(2, 3)))
        return torch.cat([self.fc_t(g), self.fc_rot(g)], dim=-1)  # (n, 7): trans | quat


class MLPHead(nn.Module):
    def __init__(self, input_feature_dim: int, output_dim: int, num_mlp_layers: int = 2, hidden_dim: Optional[int] = None,
                 activation: str = "relu", **_):
        super().__init__()
        hidden_dim = hidden_dim or input_feature_dim
        self.activation = activation
        layers: List[nn.Module] = []
        d = input_feature_dim
        for _ in range(num_mlp_layers):
            layers += [nn.Linear(d, hidden_dim), nn.ReLU() if activation == "relu" else nn.GELU()]
            d = hidden_dim
        layers.append(nn.Linear(d, output_dim))
        self.mlp = nn.Sequential(*layers)

    def forward(self, tokens: torch.Tensor) -> torch.Tensor:
        """(B, D, T) -> (B, out, T)."""
        return self.mlp(tokens.transpose(1, 2)).transpose(1, 2)


# ------------------------------------------------------------------------------------------------
# A.5 / A.6 adaptors (parameter-free)
# ------------------------------------------------------------------------------------------------
def dense_adaptor_raydirs_depth_conf_mask(x: torch.Tensor):
    """x (n, 6, H, W): [ray(3) | depth logit | confidence logit | mask logit] ->
    value (n,4,H,W) = [unit ray | exp(depth)], confidence (n,1,H,W) = 1 + exp, mask = sigmoid(logits), logits."""
    ray = x[:, 0:3]
    ray = ray / ray.norm(dim=1, keepdim=True)
    depth = torch.exp(x[:, 3:4]).clamp(min=0.0)
    conf = 1.0 + torch.exp(x[:, 4:5])
    logits = x[:, 5:6]
    return torch.cat([ray, depth], dim=1), conf, torch.sigmoid(logits), logits


# ------------------------------------------------------------------------------------------------
# The other dense adaptors MapAnything can be configured with (reference model.py:407-587; YAMLs
# configs/model/pred_head/adaptor_config/{pointmap_confidence*, campointmap_pose_*, pointmap_*raydirs_depth_pose_*}.yaml).
# [UPSTREAM-RECALL] like everything in this file: the activations follow the published DUSt3R / MoGe forms the uniception
# adaptors wrap -- "exp": unit direction x expm1(norm) (DUSt3R reg_dense_depth), "z_exp": (x*z, y*z, z) with z = exp(raw z)
# (MoGe), "linear": identity; depth "exp" = exp; confidence "exp" = vmin + exp; mask = sigmoid(logits).
# Channel order of the head output: [scene representation | confidence logit (if any) | mask logit (if any)].
# ------------------------------------------------------------------------------------------------
SCENE_REP_CHANNELS = {"pointmap": 3, "raymap+depth": 7, "raydirs+depth+pose": 4, "campointmap+pose": 3,
                      "pointmap+raydirs+depth+pose": 7}


def split_adaptor_type(adaptor_type: str):
    """'pointmap+raydirs+depth+pose+confidence+mask' -> ('pointmap+raydirs+depth+pose', has_conf, has_mask)."""
    parts = adaptor_type.split("+")
    has_mask = parts[-1] == "mask"
    if has_mask:
        parts = parts[:-1]
    has_conf = parts[-1] == "confidence"
    if has_conf:
        parts = parts[:-1]
    rep = "+".join(parts)
    if rep not in SCENE_REP_CHANNELS:
        raise ValueError(f"Invalid adaptor_type: {adaptor_type}")
    return rep, has_conf, has_mask


def point_activation(xyz: torch.Tensor, mode: str) -> torch.Tensor:
    """(n, 3, H, W) raw -> points."""
    if mode == "linear":
        return xyz
    if mode == "exp":
        d = xyz.norm(dim=1, keepdim=True)
        return xyz / d.clamp(min=1e-8) * torch.expm1(d)
    if mode == "z_exp":
        z = torch.exp(xyz[:, 2:3])
        return torch.cat([xyz[:, 0:2] * z, z], dim=1)
    raise ValueError(f"unsupported pointmap_mode {mode}")


def _unit(x: torch.Tensor) -> torch.Tensor:
    return x / x.norm(dim=1, keepdim=True)


def dense_adaptor(x: torch.Tensor, adaptor_type: str, cfg: dict):
    """x (n, C, H, W) head output -> (value (n, rep_channels, H, W), confidence | None, mask | None, logits | None)."""
    rep, has_conf, has_mask = split_adaptor_type(adaptor_type)
    c = SCENE_REP_CHANNELS[rep]
    if rep in ("pointmap", "campointmap+pose"):
        value = point_activation(x[:, 0:3], cfg.get("pointmap_mode", "exp"))
    elif rep == "raymap+depth":
        value = torch.cat([x[:, 0:3], _unit(x[:, 3:6]), torch.exp(x[:, 6:7])], dim=1)
    elif rep == "raydirs+depth+pose":
        value = torch.cat([_unit(x[:, 0:3]), torch.exp(x[:, 3:4]).clamp(min=0.0)], dim=1)
    else:  # pointmap+raydirs+depth+pose
        value = torch.cat([point_activation(x[:, 0:3], cfg.get("pointmap_mode", "exp")), _unit(x[:, 3:6]),
                           torch.exp(x[:, 6:7]).clamp(min=0.0)], dim=1)
    conf = mask = logits = None
    if has_conf:
        conf = float(cfg.get("confidence_vmin", 1)) + torch.exp(x[:, c:c + 1])
        c += 1
    if has_mask:
        logits = x[:, c:c + 1]
        mask = torch.sigmoid(logits)
    return value, conf, mask, logits


class LinearFeature(nn.Module):
    """pred_head_type 'linear' (reference model.py:339-343, :363-365): 1x1 conv D -> output_dim * patch^2, pixel shuffle."""

    def __init__(self, input_feature_dim: int, output_dim: int, patch_size: int, **_):
        super().__init__()
        self.patch_size = patch_size
        self.proj = nn.Conv2d(input_feature_dim, output_dim * patch_size * patch_size, 1)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return F.pixel_shuffle(self.proj(x), self.patch_size)


def pose_adaptor_trans_quats(x: torch.Tensor) -> torch.Tensor:
    """(n, 7) -> [trans (linear) | quats normalised]."""
    q = x[:, 3:7]
    return torch.cat([x[:, 0:3], q / q.norm(dim=1, keepdim=True)], dim=1)


def scale_adaptor_exp(x: torch.Tensor) -> torch.Tensor:
    """(B, 1, T) -> exp, clamped to [1e-8, inf)."""
    return torch.exp(x).clamp(min=1e-8)
