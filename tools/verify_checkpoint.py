#!/usr/bin/env python
"""Check a MapAnything checkpoint against the assumptions this build makes about the un-vendored `uniception` modules.

    python tools/verify_checkpoint.py <dir with config.json + model.safetensors | file.safetensors | file.pth>

SURVEY.md Appendix A lists what had to be restated from recall ("VERIFY" items).  Everything a checkpoint CAN decide is read
from its tensor names and shapes here, printed next to the value this build assumes, and turned into constructor arguments;
the model is then built (on the meta device: no memory, no GPU) and the checkpoint is loaded with strict=True semantics (every
key and shape must match).  What tensors cannot decide -- activation placement, even/odd order of global and frame blocks, the
view-PE variant when the table is not a persistent buffer -- is listed with the switch that controls it
(params.py / oracle/uniception_modules.py; tests/test_variants_gpu.py runs every value).

Exit status 0 = every key and shape matches the module tree built from the inferred configuration.
"""
from __future__ import annotations

import copy
import json
import re
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "map-anything_b200"))


def load_shapes(path: Path):
    """-> ({tensor name: shape tuple}, config dict or None)."""
    cfg = None
    if path.is_dir():
        if (path / "config.json").exists():
            cfg = json.loads((path / "config.json").read_text())
        cands = sorted(path.glob("*.safetensors")) or sorted(path.glob("*.pth")) + sorted(path.glob("*.pt"))
        if not cands:
            raise SystemExit(f"{path}: no model.safetensors / .pth found")
        path = cands[0]
    if path.suffix == ".safetensors":
        from safetensors import safe_open

        with safe_open(str(path), framework="pt") as f:
            return {k: tuple(f.get_slice(k).get_shape()) for k in f.keys()}, cfg
    import torch

    ckpt = torch.load(str(path), map_location="cpu", weights_only=False)
    sd = ckpt["model"] if isinstance(ckpt, dict) and "model" in ckpt else ckpt
    return {k: tuple(v.shape) for k, v in sd.items()}, cfg


ALIASES = (("dense_head.0.", "dpt_feature_head."), ("dense_head.1.", "dpt_regressor_head."))


def expand_aliases(shapes):
    """nn.Sequential(dense_head) re-registers the DPT modules (reference model.py:374-388): both prefixes name the same
    tensors, and safetensors stores shared tensors under ONE of the names."""
    out = dict(shapes)
    for a, b in ALIASES:
        for k, v in shapes.items():
            if k.startswith(a):
                out.setdefault(b + k[len(a):], v)
            elif k.startswith(b):
                out.setdefault(a + k[len(b):], v)
    return out


def count(shapes, pattern):
    idx = {int(m.group(1)) for k in shapes for m in [re.match(pattern, k)] if m}
    return (max(idx) + 1) if idx else 0


def infer(shapes):
    """Architecture facts readable from names + shapes: {fact: value}."""
    f = {}
    g = shapes.get
    f["encoder.embed_dim"] = g("encoder.model.cls_token", (0, 0, 0))[-1]
    f["encoder.depth"] = count(shapes, r"encoder\.model\.blocks\.(\d+)\.")
    f["encoder.patch_size"] = g("encoder.model.patch_embed.proj.weight", (0, 0, 0, 0))[-1]
    f["encoder.layer_scale"] = "encoder.model.blocks.0.ls1.gamma" in shapes
    f["encoder.register_tokens"] = g("encoder.model.register_tokens", (0, 0, 0))[1] if "encoder.model.register_tokens" in shapes else 0
    f["encoder.pos_embed_tokens"] = g("encoder.model.pos_embed", (0, 0, 0))[1]
    f["info.dim"] = g("info_sharing.norm.weight", (0,))[0]
    f["info.proj_embed"] = g("info_sharing.proj_embed.weight")
    f["info.depth"] = count(shapes, r"info_sharing\.self_attention_blocks\.(\d+)\.")
    fc1 = g("info_sharing.self_attention_blocks.0.mlp.fc1.weight")
    f["info.mlp_ratio"] = (fc1[0] / fc1[1]) if fc1 else None
    f["info.layer_scale"] = any(k.startswith("info_sharing.self_attention_blocks.0.ls1") for k in shapes)
    f["info.qk_norm"] = any(re.match(r"info_sharing\.self_attention_blocks\.0\.attn\.(q_norm|k_norm)\.", k) for k in shapes)
    f["info.qkv_bias"] = "info_sharing.self_attention_blocks.0.attn.qkv.bias" in shapes
    f["info.persistent_view_pe"] = [k for k in shapes if k.startswith("info_sharing.") and re.search(r"pos|pe_|table", k)]
    f["info.other_block_types"] = sorted({k.split(".")[1] for k in shapes if k.startswith("info_sharing.")}
                                         - {"proj_embed", "self_attention_blocks", "norm"})
    f["dpt.feature_dim"] = g("dpt_feature_head.scratch.layer_rn.0.weight", (0,))[0]
    f["dpt.layer_dims"] = [g(f"dpt_feature_head.scratch.layer_rn.{i}.weight", (0, 0))[1] for i in range(4)]
    f["dpt.input_feature_dims"] = [g(f"dpt_feature_head.act_postprocess.{i}.0.weight", (0, 0))[1] for i in range(4)]
    f["dpt.layer_rn_bias"] = "dpt_feature_head.scratch.layer_rn.0.bias" in shapes
    f["dpt.refinenet4_has_rcu1"] = "dpt_feature_head.scratch.refinenet4.resConfUnit1.conv1.weight" in shapes
    f["regressor.hidden_dims"] = [g("dpt_regressor_head.conv1.weight", (0,))[0], g("dpt_regressor_head.conv2.0.weight", (0,))[0]]
    f["regressor.output_dim"] = g("dpt_regressor_head.conv2.2.weight", (0,))[0]
    f["pose.num_resconv_block"] = count(shapes, r"pose_head\.res_conv\.(\d+)\.")
    f["pose.skip_projection"] = "pose_head.res_conv.0.head_skip.weight" in shapes
    f["pose.more_mlps_linears"] = len({k.split(".")[2] for k in shapes if k.startswith("pose_head.more_mlps.") and k.endswith("weight")})
    f["pose.rot_dim"] = g("pose_head.fc_rot.weight", (0,))[0]
    lin = sorted({int(k.split(".")[2]) for k in shapes if re.match(r"scale_head\.mlp\.\d+\.weight", k)})
    f["scale.linear_indices"] = lin
    f["scale.widths"] = [g(f"scale_head.mlp.{i}.weight") for i in lin]
    for enc in ("ray_dirs_encoder", "depth_encoder"):
        f[f"{enc}.conv_in"] = g(f"{enc}.conv_in.weight")
        f[f"{enc}.dims"] = [g(f"{enc}.encoder.0.conv1.weight", (0,))[0], g(f"{enc}.encoder.1.conv1.weight", (0,))[0]]
    for enc in ("depth_scale_encoder", "cam_rot_encoder", "cam_trans_encoder", "cam_trans_scale_encoder"):
        idx = sorted({int(k.split(".")[2]) for k in shapes if re.match(rf"{enc}\.encoder\.\d+\.weight", k)})
        f[f"{enc}.dims"] = [g(f"{enc}.encoder.{i}.weight", (0,))[0] for i in idx]
    f["dense_head_aliases"] = any(k.startswith("dense_head.") for k in shapes)
    # prediction head type (reference model.py:339-388) and the adaptor types its channel count admits (model.py:407-587)
    linear = "dense_head.proj.weight" in shapes
    posed = any(k.startswith("pose_head.") for k in shapes)
    f["head.type"] = "linear" if linear else ("dpt+pose" if posed else "dpt")
    if linear:
        p2 = max(1, f["encoder.patch_size"]) ** 2
        f["head.output_dim"] = g("dense_head.proj.weight", (0,))[0] // p2
    else:
        f["head.output_dim"] = f["regressor.output_dim"]
    f["head.adaptor_candidates"] = adaptor_candidates(f["head.output_dim"], posed)
    return f


REP_CHANNELS = {"pointmap": 3, "raymap+depth": 7, "raydirs+depth+pose": 4, "campointmap+pose": 3, "pointmap+raydirs+depth+pose": 7}


def adaptor_candidates(channels, posed):
    """adaptor types whose channel count (representation + confidence + mask) equals the dense head's output width; the posed
    representations need the pose head.  The released type first."""
    out = []
    for rep, c in REP_CHANNELS.items():
        if ("pose" in rep) != posed:
            continue
        for suffix, extra in (("+confidence+mask", 2), ("+confidence", 1), ("+mask", 1), ("", 0)):
            if c + extra == channels:
                out.append(rep + suffix)
    return sorted(out, key=lambda t: t != "raydirs+depth+pose+confidence+mask")


def config_from(facts, cfg_json):
    """Constructor kwargs: the checkpoint's config.json when present (the reference's HF layout), else this build's released
    configuration adjusted by what the tensors say."""
    from mapanything_b200.config import mapanything_config

    if cfg_json is not None and "encoder_config" in cfg_json:
        keys = ("name", "encoder_config", "info_sharing_config", "pred_head_config", "geometric_input_config")
        return {k: copy.deepcopy(cfg_json[k]) for k in keys}
    cfg = mapanything_config()
    ma = cfg["info_sharing_config"]["module_args"]
    ma.update({"dim": facts["info.dim"], "num_heads": facts["info.dim"] // 64, "depth": facts["info.depth"]})
    if facts["info.depth"] == 48:
        ma["indices"] = [11, 23, 35]
    cfg["encoder_config"]["vit_kwargs"] = {"embed_dim": facts["encoder.embed_dim"], "depth": facts["encoder.depth"],
                                           "num_heads": facts["encoder.embed_dim"] // 64,
                                           "img_size": int(round((facts["encoder.pos_embed_tokens"] - 1) ** 0.5)) * facts["encoder.patch_size"]}
    cand = facts["head.adaptor_candidates"]
    if facts["head.type"] != "dpt+pose" or (cand and cand[0] != "raydirs+depth+pose+confidence+mask"):
        # another head / scene representation: the first adaptor type the channel count admits (tensors cannot tell e.g.
        # "+confidence" from "+mask"; both build the same module tree)
        from mapanything_b200.config import ADAPTOR_CONFIGS, pred_head_variant_config

        if not cand:
            raise SystemExit(f"no adaptor type consumes {facts['head.output_dim']} channels with head {facts['head.type']}")
        rep = cand[0].replace("+confidence", "").replace("+mask", "")
        yaml = next(k for k, v in ADAPTOR_CONFIGS.items() if v["scene_rep_type"] == rep) if rep != "raymap+depth" else None
        if yaml is not None:
            cfg["pred_head_config"] = pred_head_variant_config(yaml, head_type=facts["head.type"], adaptor_type=cand[0])
        else:
            ph = cfg["pred_head_config"]
            ph.update({"type": facts["head.type"], "adaptor_type": cand[0], "adaptor": {"name": cand[0]}})
            ph["regressor_head"]["output_dim"] = facts["head.output_dim"]
    if facts["head.type"] != "linear":
        cfg["pred_head_config"]["regressor_head"]["hidden_dims"] = facts["regressor.hidden_dims"]
    if facts["head.type"] == "dpt+pose":
        cfg["pred_head_config"]["pose_head"]["num_resconv_block"] = facts["pose.num_resconv_block"]
    n_lin = len(facts["scale.linear_indices"])
    if n_lin:
        cfg["pred_head_config"]["scale_head"].update({"num_mlp_layers": n_lin - 1, "hidden_dim": facts["scale.widths"][0][0]})
    return cfg


ASSUMED = {
    "encoder.layer_scale": True, "encoder.register_tokens": 0, "info.mlp_ratio": 4.0, "info.layer_scale": False,
    "info.qk_norm": False, "info.qkv_bias": True, "info.persistent_view_pe": [], "info.other_block_types": [],
    "dpt.feature_dim": 256, "dpt.layer_dims": [96, 192, 384, 768], "dpt.layer_rn_bias": False, "dpt.refinenet4_has_rcu1": True,
    "regressor.hidden_dims": [128, 128], "regressor.output_dim": 6, "pose.num_resconv_block": 2, "pose.skip_projection": False,
    "pose.more_mlps_linears": 2, "pose.rot_dim": 4, "head.type": "dpt+pose", "ray_dirs_encoder.dims": [768, 1024], "depth_encoder.dims": [768, 1024],
}

UNDECIDABLE = [
    ("even blocks global vs frame-wise", "info_sharing module_args global_attention_first (default True)"),
    ("view PE: row 0 for view 0 only / rows 0 and 1 / per-view index", "view_pe_variant (default 'ref_only'), unless a table is listed above"),
    ("ResConvBlock: relu(skip + conv3) vs skip + relu(conv3)", "pose_head final_relu_after_skip (default True)"),
    ("MLPHead activation", "scale_head activation (default 'relu')"),
    ("depth adaptor exp vs expm1, confidence 1 + exp", "adaptor config: depth_mode / confidence_type (only 'exp' implemented)"),
    ("which of head.adaptor_candidates the model was trained with", "pred_head_config adaptor_type (config.json carries it)"),
    ("point activation of pointmap / campointmap representations", "adaptor pointmap_mode: 'exp' | 'z_exp' | 'linear'"),
    ("entropy-scaling constant", "entropy_scaling_ref_len (only used when use_entropy_scaling)"),
    ("GELU (erf) in MLPs, LayerNorm eps 1e-6", "fixed; DINOv2 / timm convention"),
]


def main(argv):
    if len(argv) != 2:
        print(__doc__)
        return 2
    shapes, cfg_json = load_shapes(Path(argv[1]))
    shapes = expand_aliases(shapes)
    facts = infer(shapes)
    print(f"{len(shapes)} tensors" + ("" if cfg_json is None else " + config.json"))
    print("\n== read from tensor names / shapes (assumed by the released config -> found)")
    bad = 0
    for k, v in facts.items():
        tag = ""
        if k in ASSUMED:
            ok = ASSUMED[k] == v
            tag = "  ok" if ok else f"  DIFFERS (this build assumes {ASSUMED[k]})"
            bad += 0 if ok else 1
        print(f"  {k:34s} {v}{tag}")
    print("\n== not decidable from tensors (switch that controls it)")
    for what, switch in UNDECIDABLE:
        print(f"  {what:62s} -> {switch}")

    import torch

    from mapanything_b200 import MapAnything

    cfg = config_from(facts, cfg_json)
    with torch.device("meta"):
        model = MapAnything(**cfg)
    ours = {k: tuple(v.shape) for k, v in model.state_dict().items()}
    missing = sorted(set(ours) - set(shapes))
    unexpected = sorted(set(shapes) - set(ours))
    mismatched = sorted(k for k in set(ours) & set(shapes) if ours[k] != shapes[k])
    print("\n== strict load against the module tree built from that configuration")
    for name, lst in (("missing (module has, checkpoint lacks)", missing), ("unexpected (checkpoint has, module lacks)", unexpected),
                      ("shape mismatch", mismatched)):
        print(f"  {name}: {len(lst)}")
        for k in lst[:20]:
            print(f"     {k}  ours {ours.get(k)}  checkpoint {shapes.get(k)}")
    ok = not (missing or unexpected or mismatched)
    print("\nRESULT:", "strict load OK -- every key and shape matches" if ok else "MISMATCH -- see above")
    if bad:
        print(f"        {bad} architecture fact(s) differ from the released configuration (handled by the inferred config if the load is OK)")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main(sys.argv))
