"""Parameter tree of the drop-in MapAnything module.

torch.nn modules are used here ONLY as named parameter containers so that `state_dict()` /
`load_state_dict()` / `.to()` / `named_parameters()` behave like the reference's module (same key prefixes:
reference model.py:157-202, :299, :374-388; sub-keys per SURVEY.md App. A.7).  None of their `forward`
methods is ever called: all arithmetic runs in the CUDA kernels driven by `engine.py`.
"""
from __future__ import annotations

from typing import Optional, Sequence

import torch
import torch.nn as nn


class _NoForward(nn.Module):
    def forward(self, *a, **k):  # pragma: no cover - guards against an accidental PyTorch compute path
        raise RuntimeError(f"{type(self).__name__} is a parameter container; compute runs in libmapanything_b200.so")


class LayerScale(_NoForward):
    def __init__(self, dim):
        super().__init__()
        self.gamma = nn.Parameter(torch.ones(dim))


class Attention(_NoForward):
    def __init__(self, dim, heads):
        super().__init__()
        self.num_heads = heads
        self.qkv = nn.Linear(dim, 3 * dim)
        self.proj = nn.Linear(dim, dim)


class Mlp(_NoForward):
    def __init__(self, dim, hidden):
        super().__init__()
        self.fc1 = nn.Linear(dim, hidden)
        self.fc2 = nn.Linear(hidden, dim)


class Block(_NoForward):
    def __init__(self, dim, heads, mlp_ratio=4.0, layer_scale=True):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim, eps=1e-6)
        self.attn = Attention(dim, heads)
        self.ls1 = LayerScale(dim) if layer_scale else nn.Identity()
        self.norm2 = nn.LayerNorm(dim, eps=1e-6)
        self.mlp = Mlp(dim, int(dim * mlp_ratio))
        self.ls2 = LayerScale(dim) if layer_scale else nn.Identity()


class PatchEmbed(_NoForward):
    def __init__(self, patch, dim):
        super().__init__()
        self.proj = nn.Conv2d(3, dim, patch, stride=patch)


class DinoV2(_NoForward):
    def __init__(self, img_size=518, patch_size=14, embed_dim=1024, depth=24, num_heads=16, mlp_ratio=4.0,
                 interpolate_offset=0.1):
        super().__init__()
        self.patch_size, self.embed_dim, self.num_heads = patch_size, embed_dim, num_heads
        self.interpolate_offset = interpolate_offset
        n = (img_size // patch_size) ** 2
        self.cls_token = nn.Parameter(torch.zeros(1, 1, embed_dim))
        self.pos_embed = nn.Parameter(torch.zeros(1, n + 1, embed_dim))
        self.mask_token = nn.Parameter(torch.zeros(1, embed_dim))
        self.patch_embed = PatchEmbed(patch_size, embed_dim)
        self.blocks = nn.ModuleList([Block(embed_dim, num_heads, mlp_ratio) for _ in range(depth)])
        self.norm = nn.LayerNorm(embed_dim, eps=1e-6)


class DinoV2Encoder(_NoForward):
    def __init__(self, name="dinov2_large", data_norm_type="dinov2", size="large", with_registers=False,
                 gradient_checkpointing=False, torch_hub_force_reload=False, vit_kwargs=None, encoder_str="dinov2", **_):
        super().__init__()
        if size != "large" or with_registers:
            raise ValueError("mapanything_b200 implements the ViT-L/14 (no registers) encoder of the released model")
        self.name, self.data_norm_type = name, data_norm_type
        self.model = DinoV2(**(vit_kwargs or {}))
        self.patch_size, self.enc_embed_dim = self.model.patch_size, self.model.embed_dim


class ResidualBlock(_NoForward):
    def __init__(self, cin, cout):
        super().__init__()
        self.conv1 = nn.Conv2d(cin, cout, 3, padding=1)
        self.conv2 = nn.Conv2d(cout, cout, 3, padding=1)
        self.shortcut = nn.Identity() if cin == cout else nn.Conv2d(cin, cout, 1)


class DenseRepresentationEncoder(_NoForward):
    def __init__(self, name, in_chans, enc_embed_dim, patch_size, apply_pe=False, intermediate_dims=(588, 768, 1024),
                 encoder_str="dense_rep_encoder", **_):
        super().__init__()
        if apply_pe:
            raise ValueError("apply_pe=True is not used by any MapAnything task config")
        self.name, self.patch_size, self.enc_embed_dim, self.in_chans = name, patch_size, enc_embed_dim, in_chans
        d0, d1, d2 = intermediate_dims
        self.conv_in = nn.Conv2d(in_chans * patch_size * patch_size, d0, 3, padding=1)
        self.encoder = nn.Sequential(ResidualBlock(d0, d1), ResidualBlock(d1, d2), nn.Conv2d(d2, enc_embed_dim, 1))
        self.norm_layer = nn.LayerNorm(enc_embed_dim, eps=1e-6)


class GlobalRepresentationEncoder(_NoForward):
    def __init__(self, name, in_chans, enc_embed_dim, intermediate_dims=(128, 256, 512), encoder_str="global_rep_encoder", **_):
        super().__init__()
        self.name, self.enc_embed_dim, self.in_chans = name, enc_embed_dim, in_chans
        dims = [in_chans, *intermediate_dims, enc_embed_dim]
        layers = []
        for i in range(len(dims) - 1):
            layers.append(nn.Linear(dims[i], dims[i + 1]))
            if i < len(dims) - 2:
                layers.append(nn.GELU())
        self.encoder = nn.Sequential(*layers)
        self.norm_layer = nn.LayerNorm(enc_embed_dim, eps=1e-6)


class AlternatingAttentionIFR(_NoForward):
    """Multi-view transformer with intermediate-feature return.  Besides the YAML keys of
    configs/model/info_sharing/*.yaml it takes the switches that settle the App. A.3 "VERIFY" items (the oracle has the same
    ones; tests/test_variants_gpu.py runs each):
      global_attention_first  even blocks global (default) or frame-wise
      attention_pattern       "alternating" | "global" (model_type global_attention: every block attends over all views)
      view_pe_variant         "ref_only" | "ref_vs_rest" | "per_view_index" (+ use_rand_idx_pe_for_non_reference_views)
      use_entropy_scaling     softmax scale x log(n_keys) / log(entropy_scaling_ref_len)"""

    def __init__(self, name, input_embed_dim, indices=(11, 17), norm_intermediate=True, size=None, depth=24, dim=768,
                 num_heads=12, mlp_ratio=4.0, distinguish_ref_and_non_ref_views=True, use_pe_for_non_reference_views=False,
                 use_rand_idx_pe_for_non_reference_views=False, max_num_views_for_pe=1000, max_num_views=None,
                 gradient_checkpointing=False, custom_positional_encoding=None, global_attention_first=True,
                 attention_pattern="alternating", view_pe_variant=None, use_entropy_scaling=False,
                 entropy_scaling_ref_len=None, **_):
        super().__init__()
        if custom_positional_encoding is not None:
            raise ValueError("custom positional encodings are not implemented by the reference either (model.py:262-268)")
        if dim // num_heads != 64:
            raise ValueError("the sm_100a attention kernel is specialised for head_dim 64")
        if attention_pattern not in ("alternating", "global"):
            raise ValueError(f"attention_pattern {attention_pattern!r}: 'alternating' or 'global'")
        if view_pe_variant is None:
            view_pe_variant = "per_view_index" if (use_rand_idx_pe_for_non_reference_views or use_pe_for_non_reference_views) \
                else "ref_only"
        if view_pe_variant not in ("ref_only", "ref_vs_rest", "per_view_index"):
            raise ValueError(f"view_pe_variant {view_pe_variant!r}")
        self.name, self.dim, self.depth, self.num_heads = name, dim, depth, num_heads
        self.indices, self.norm_intermediate = list(indices), norm_intermediate
        self.distinguish_ref_and_non_ref_views = distinguish_ref_and_non_ref_views
        self.global_attention_first, self.attention_pattern = bool(global_attention_first), attention_pattern
        self.view_pe_variant, self.use_rand_idx = view_pe_variant, bool(use_rand_idx_pe_for_non_reference_views)
        self.max_views = int(max_num_views or max_num_views_for_pe)
        self.fixed_view_pe_indices = None  # tests / reproducible runs: table rows of views 1..V-1 instead of random draws
        self.use_entropy_scaling, self.entropy_scaling_ref_len = bool(use_entropy_scaling), entropy_scaling_ref_len
        self.proj_embed = nn.Linear(input_embed_dim, dim) if input_embed_dim != dim else nn.Identity()
        self.self_attention_blocks = nn.ModuleList([Block(dim, num_heads, mlp_ratio, layer_scale=False) for _ in range(depth)])
        self.norm = nn.LayerNorm(dim, eps=1e-6)

    def is_global(self, i: int) -> bool:
        return self.attention_pattern == "global" or ((i % 2 == 0) == self.global_attention_first)

    def view_pe_rows(self, v: int, first_view: int = 0):
        """Sinusoid-table row per view (None = no PE) for the scene's views first_view .. first_view + v - 1."""
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
            elif self.fixed_view_pe_indices is not None:
                rows.append(int(self.fixed_view_pe_indices[g - 1]))
            elif self.use_rand_idx:
                rows.append(int(torch.randint(1, self.max_views, (1,))))
            else:
                rows.append(g)
        return rows

    def softmax_scale_factor(self, n_keys: int) -> float:
        if not self.use_entropy_scaling:
            return 1.0
        import math

        return math.log(n_keys) / math.log(self.entropy_scaling_ref_len or n_keys)


class ResidualConvUnit(_NoForward):
    def __init__(self, ch):
        super().__init__()
        self.conv1 = nn.Conv2d(ch, ch, 3, padding=1)
        self.conv2 = nn.Conv2d(ch, ch, 3, padding=1)


class FeatureFusionBlock(_NoForward):
    def __init__(self, ch):
        super().__init__()
        self.resConfUnit1 = ResidualConvUnit(ch)
        self.resConfUnit2 = ResidualConvUnit(ch)
        self.out_conv = nn.Conv2d(ch, ch, 1)


class Scratch(_NoForward):
    def __init__(self, layer_dims, feature_dim):
        super().__init__()
        self.layer_rn = nn.ModuleList([nn.Conv2d(d, feature_dim, 3, padding=1, bias=False) for d in layer_dims])
        self.refinenet1 = FeatureFusionBlock(feature_dim)
        self.refinenet2 = FeatureFusionBlock(feature_dim)
        self.refinenet3 = FeatureFusionBlock(feature_dim)
        self.refinenet4 = FeatureFusionBlock(feature_dim)


class DPTFeature(_NoForward):
    def __init__(self, patch_size, input_feature_dims: Sequence[int], feature_dim=256, hooks=(0, 1, 2, 3),
                 layer_dims=(96, 192, 384, 768), **_):
        super().__init__()
        self.hooks, self.feature_dim, self.layer_dims = list(hooks), feature_dim, list(layer_dims)
        d, ld = list(input_feature_dims), list(layer_dims)
        self.act_postprocess = nn.ModuleList(
            [
                nn.Sequential(nn.Conv2d(d[0], ld[0], 1), nn.ConvTranspose2d(ld[0], ld[0], 4, stride=4)),
                nn.Sequential(nn.Conv2d(d[1], ld[1], 1), nn.ConvTranspose2d(ld[1], ld[1], 2, stride=2)),
                nn.Sequential(nn.Conv2d(d[2], ld[2], 1)),
                nn.Sequential(nn.Conv2d(d[3], ld[3], 1), nn.Conv2d(ld[3], ld[3], 3, stride=2, padding=1)),
            ]
        )
        self.scratch = Scratch(ld, feature_dim)


class DPTRegressionProcessor(_NoForward):
    def __init__(self, input_feature_dim, output_dim, hidden_dims: Optional[Sequence[int]] = None, **_):
        super().__init__()
        hidden_dims = list(hidden_dims) if hidden_dims is not None else [input_feature_dim // 2, input_feature_dim // 2]
        self.conv1 = nn.Conv2d(input_feature_dim, hidden_dims[0], 3, padding=1)
        self.conv2 = nn.Sequential(nn.Conv2d(hidden_dims[0], hidden_dims[1], 3, padding=1), nn.ReLU(),
                                   nn.Conv2d(hidden_dims[1], output_dim, 1))


class LinearFeature(_NoForward):
    """pred_head_type "linear" (reference model.py:339-343, :363-365): 1x1 conv D -> output_dim * patch^2, then pixel shuffle."""

    def __init__(self, input_feature_dim, output_dim, patch_size, **_):
        super().__init__()
        self.output_dim, self.patch_size = output_dim, patch_size
        self.proj = nn.Conv2d(input_feature_dim, output_dim * patch_size * patch_size, 1)


class ResConvBlock(_NoForward):
    def __init__(self, cin, cout):
        super().__init__()
        self.head_skip = nn.Identity() if cin == cout else nn.Conv2d(cin, cout, 1)
        self.res_conv1 = nn.Conv2d(cin, cout, 1)
        self.res_conv2 = nn.Conv2d(cout, cout, 3, padding=1)
        self.res_conv3 = nn.Conv2d(cout, cout, 1)


class PoseHead(_NoForward):
    """final_relu_after_skip: relu(skip + conv3(y)) (default) or skip + relu(conv3(y)) -- SURVEY App. A.6 VERIFY item."""

    def __init__(self, patch_size, input_feature_dim, num_resconv_block=2, rot_representation_dim=4,
                 final_relu_after_skip=True, **_):
        super().__init__()
        self.final_relu_after_skip = bool(final_relu_after_skip)
        c = input_feature_dim
        self.res_conv = nn.ModuleList([ResConvBlock(c, c) for _ in range(num_resconv_block)])
        self.more_mlps = nn.Sequential(nn.Linear(c, c), nn.ReLU(), nn.Linear(c, c), nn.ReLU())
        self.fc_t = nn.Linear(c, 3)
        self.fc_rot = nn.Linear(c, rot_representation_dim)


class MLPHead(_NoForward):
    """num_mlp_layers / hidden_dim / activation ("relu" | "gelu"): SURVEY App. A.6 VERIFY items."""

    def __init__(self, input_feature_dim, output_dim, num_mlp_layers=2, hidden_dim=None, activation="relu", **_):
        super().__init__()
        if activation not in ("relu", "gelu"):
            raise ValueError(f"MLPHead activation {activation!r}")
        hidden_dim = hidden_dim or input_feature_dim
        self.activation = activation
        layers, d = [], input_feature_dim
        for _i in range(num_mlp_layers):
            layers += [nn.Linear(d, hidden_dim), nn.ReLU() if activation == "relu" else nn.GELU()]
            d = hidden_dim
        layers.append(nn.Linear(d, output_dim))
        self.mlp = nn.Sequential(*layers)


def encoder_factory(encoder_str: str, **kw) -> nn.Module:
    """Mirror of uniception's `encoder_factory` for the three encoder kinds MapAnything instantiates (model.py:157-193)."""
    if encoder_str == "dinov2":
        return DinoV2Encoder(**kw)
    if encoder_str == "dense_rep_encoder":
        return DenseRepresentationEncoder(**kw)
    if encoder_str == "global_rep_encoder":
        return GlobalRepresentationEncoder(**kw)
    raise ValueError(f"Unknown encoder_str: {encoder_str}")
