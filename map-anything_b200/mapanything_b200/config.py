"""The reference's Hydra composition for the released model, as plain dicts (what `configs/model/mapanything.yaml` composes to).

Mirrors /root/reference/configs/model/mapanything.yaml:1-18 with
  encoder:      configs/model/encoder/dinov2_large.yaml
  info_sharing: configs/model/info_sharing/aat_ifr_24_layers.yaml
  pred_head:    configs/model/pred_head/dpt_pose_scale.yaml (+ adaptor_config/raydirs_depth_pose_confidence_mask_scale.yaml)
  task:         configs/model/task/images_only.yaml (defaults: task/default.yaml)
"""
from __future__ import annotations

import copy

INF = float("inf")


def mapanything_config(**overrides) -> dict:
    cfg = {
        "name": "mapanything",
        "encoder_config": {
            "encoder_str": "dinov2", "name": "dinov2_large", "data_norm_type": "dinov2", "size": "large",
            "with_registers": False, "uses_torch_hub": True, "gradient_checkpointing": False,
        },
        "info_sharing_config": {
            "model_type": "alternating_attention", "model_return_type": "intermediate_features",
            "custom_positional_encoding": None,
            "module_args": {
                "name": "aat_24_layers_ifr", "indices": [11, 17], "norm_intermediate": True, "size": "24_layers",
                "depth": 24, "distinguish_ref_and_non_ref_views": True, "gradient_checkpointing": False,
            },
        },
        "pred_head_config": {
            "type": "dpt+pose",
            "feature_head": {"feature_dim": 256, "hooks": [0, 1, 2, 3], "checkpoint_gradient": False},
            "regressor_head": {"output_dim": 6, "checkpoint_gradient": False},
            "pose_head": {"num_resconv_block": 2, "rot_representation_dim": 4},
            "scale_head": {"output_dim": 1},
            "adaptor_type": "raydirs+depth+pose+confidence+mask",
            "dpt_adaptor": {
                "name": "raydirs+depth+pose+confidence+mask+scale", "ray_directions_mode": "linear",
                "ray_directions_normalize_to_unit_sphere": True, "ray_directions_normalize_to_unit_image_plane": False,
                "ray_directions_vmin": -INF, "ray_directions_vmax": INF, "ray_directions_clamp_min_of_z_dir": False,
                "ray_directions_z_dir_min": -INF, "depth_mode": "exp", "depth_vmin": 0, "depth_vmax": INF,
                "confidence_type": "exp", "confidence_vmin": 1, "confidence_vmax": INF,
            },
            "pose_adaptor": {
                "name": "raydirs+depth+pose+confidence+mask+scale", "cam_trans_mode": "linear", "cam_trans_vmin": -INF,
                "cam_trans_vmax": INF, "quaternions_mode": "linear", "quaternions_normalize": True,
                "quaternions_vmin": -INF, "quaternions_vmax": INF,
            },
            "scale_adaptor": {"name": "raydirs+depth+pose+confidence+mask+scale", "mode": "exp", "vmin": 1e-08, "vmax": INF},
            "gradient_checkpointing": False,
        },
        "geometric_input_config": {
            "ray_dirs_encoder_config": {"name": "ray_dirs_encoder", "in_chans": 3, "encoder_str": "dense_rep_encoder", "apply_pe": False},
            "depth_encoder_config": {"name": "depth_encoder", "in_chans": 1, "encoder_str": "dense_rep_encoder", "apply_pe": False},
            "cam_rot_encoder_config": {"name": "cam_rot_quats_encoder", "in_chans": 4, "encoder_str": "global_rep_encoder"},
            "cam_trans_encoder_config": {"name": "cam_trans_encoder", "in_chans": 3, "encoder_str": "global_rep_encoder"},
            "scale_encoder_config": {"name": "scale_encoder", "in_chans": 1, "encoder_str": "global_rep_encoder"},
            "overall_prob": 0, "dropout_prob": 1, "ray_dirs_prob": 0, "depth_prob": 0, "cam_prob": 0,
            "sparse_depth_prob": 0, "sparsification_removal_percent": 0, "depth_scale_norm_all_prob": 0,
            "pose_scale_norm_all_prob": 0,
        },
    }
    cfg = copy.deepcopy(cfg)
    for k, v in overrides.items():
        cfg[k] = v
    return cfg


# module_args of the alternating-attention variants under the reference's configs/model/info_sharing/
INFO_SHARING_VARIANTS = {
    "aat_ifr_24_layers": {"name": "aat_24_layers_ifr", "indices": [11, 17], "size": "24_layers", "depth": 24,
                          "distinguish_ref_and_non_ref_views": True},
    "aat_ifr_24_layers_no_ref_view": {"name": "aat_24_layers_ifr_no_ref_view", "indices": [11, 17], "size": "24_layers",
                                      "depth": 24, "distinguish_ref_and_non_ref_views": False},
    "aat_ifr_48_layers": {"name": "aat_48_layers_ifr", "indices": [11, 23, 35], "size": "48_layers", "depth": 48, "dim": 1024,
                          "num_heads": 16, "distinguish_ref_and_non_ref_views": True},
    "aat_ifr_48_layers_no_ref_view": {"name": "aat_48_layers_ifr_no_ref_view", "indices": [11, 23, 35], "size": "48_layers",
                                      "depth": 48, "dim": 1024, "num_heads": 16, "distinguish_ref_and_non_ref_views": False},
    # view-index positional encoding: random table rows for the non-reference views (aat_ifr_24_layers_w_view_pe.yaml)
    "aat_ifr_24_layers_w_view_pe": {"name": "aat_24_layers_ifr_w_view_pe", "indices": [11, 17], "size": "24_layers", "depth": 24,
                                    "distinguish_ref_and_non_ref_views": True, "max_num_views_for_pe": 1000,
                                    "use_rand_idx_pe_for_non_reference_views": True},
    # entropy scaling of the attention logits (aat_ifr_24_layers_escaling.yaml / aat_ifr_48_layers_escaling.yaml)
    "aat_ifr_24_layers_escaling": {"name": "aat_24_layers_ifr", "indices": [11, 17], "size": "24_layers", "depth": 24,
                                   "distinguish_ref_and_non_ref_views": True, "use_entropy_scaling": True},
    "aat_ifr_48_layers_escaling": {"name": "aat_48_layers_ifr", "indices": [11, 23, 35], "size": "48_layers", "depth": 48,
                                   "dim": 1024, "num_heads": 16, "distinguish_ref_and_non_ref_views": True,
                                   "use_entropy_scaling": True},
    # global attention in every block (model_type "global_attention": gat_ifr_24_layers.yaml, gat_ifr_24_layers_escaling.yaml)
    "gat_ifr_24_layers": {"name": "gat_24_layers_ifr", "indices": [11, 17], "size": "24_layers", "depth": 24,
                          "max_num_views": 1000, "use_rand_idx_pe_for_non_reference_views": True},
    "gat_ifr_24_layers_escaling": {"name": "gat_24_layers_ifr", "indices": [11, 17], "size": "24_layers", "depth": 24,
                                   "max_num_views": 1000, "use_rand_idx_pe_for_non_reference_views": True,
                                   "use_entropy_scaling": True},
}


def mapanything_variant_config(info_sharing: str = "aat_ifr_24_layers", adaptor_config: str = None, head_type: str = None,
                               adaptor_type: str = None, **overrides) -> dict:
    """mapanything_config() with another info-sharing YAML of the reference (`model/info_sharing=<name>` on its Hydra
    command line): the 48-layer / width-1024 transformer with three taps, and the variants without reference-view embedding."""
    if info_sharing not in INFO_SHARING_VARIANTS:
        raise ValueError(f"info_sharing must be one of {sorted(INFO_SHARING_VARIANTS)} (intermediate-feature transformers with "
                         f"alternating or global attention), got {info_sharing!r}")
    cfg = mapanything_config(**overrides)
    cfg["info_sharing_config"]["module_args"] = {"norm_intermediate": True, "gradient_checkpointing": False,
                                                 **copy.deepcopy(INFO_SHARING_VARIANTS[info_sharing])}
    if info_sharing.startswith("gat_"):
        cfg["info_sharing_config"]["model_type"] = "global_attention"
    if adaptor_config is not None or head_type is not None or adaptor_type is not None:
        cfg["pred_head_config"] = pred_head_variant_config(adaptor_config or "raydirs_depth_pose_confidence_mask_scale",
                                                           head_type, adaptor_type)
    return cfg


# configs/model/pred_head/adaptor_config/*.yaml of the reference (the *_scale forms: every working MapAnything config
# carries a scale head, model.py:352,388), keyed by file name.  (rep channels, point mode, posed, extra keys)
_DENSE_RAYS_DEPTH = {
    "ray_directions_mode": "linear", "ray_directions_normalize_to_unit_sphere": True,
    "ray_directions_normalize_to_unit_image_plane": False, "ray_directions_vmin": -INF, "ray_directions_vmax": INF,
    "ray_directions_clamp_min_of_z_dir": False, "ray_directions_z_dir_min": -INF, "depth_mode": "exp", "depth_vmin": 0,
    "depth_vmax": INF,
}
_CONF = {"confidence_type": "exp", "confidence_vmin": 1, "confidence_vmax": INF}
ADAPTOR_CONFIGS = {
    "raydirs_depth_pose_confidence_mask_scale": {
        "input_dim": 6, "scene_rep_dim": 4, "type": "raydirs+depth+pose+confidence+mask", "scene_rep_type": "raydirs+depth+pose",
        "dense": {**_DENSE_RAYS_DEPTH, **_CONF}},
    "pointmap_confidence_mask_scale": {
        "input_dim": 5, "scene_rep_dim": 3, "type": "pointmap+confidence+mask", "scene_rep_type": "pointmap",
        "dense": {"pointmap_mode": "exp", "pointmap_vmin": -INF, "pointmap_vmax": INF, **_CONF}},
    "campointmap_pose_confidence_mask_scale": {
        "input_dim": 5, "scene_rep_dim": 3, "type": "campointmap+pose+confidence+mask", "scene_rep_type": "campointmap+pose",
        "dense": {"pointmap_mode": "z_exp", "pointmap_vmin": -INF, "pointmap_vmax": INF, **_CONF}},
    "pointmap_raydirs_depth_pose_confidence_mask_scale": {
        "input_dim": 9, "scene_rep_dim": 7, "type": "pointmap+raydirs+depth+pose+confidence+mask",
        "scene_rep_type": "pointmap+raydirs+depth+pose", "use_factored_predictions_for_global_pointmaps": False,
        "dense": {"pointmap_mode": "exp", "pointmap_vmin": -INF, "pointmap_vmax": INF, **_DENSE_RAYS_DEPTH, **_CONF}},
    "pointmap_factored_raydirs_depth_pose_confidence_mask_scale": {
        "input_dim": 9, "scene_rep_dim": 7, "type": "pointmap+raydirs+depth+pose+confidence+mask",
        "scene_rep_type": "pointmap+raydirs+depth+pose", "use_factored_predictions_for_global_pointmaps": True,
        "dense": {"pointmap_mode": "exp", "pointmap_vmin": -INF, "pointmap_vmax": INF, **_DENSE_RAYS_DEPTH, **_CONF}},
}


def pred_head_variant_config(adaptor_config: str = "raydirs_depth_pose_confidence_mask_scale", head_type: str = None,
                             adaptor_type: str = None) -> dict:
    """`pred_head_config` for another adaptor YAML of the reference (Hydra: `model/pred_head=dpt_pose_scale
    model/pred_head/adaptor_config=<name>`, or `model/pred_head=dpt_scale` for the pose-free representations).
    head_type defaults to "dpt+pose" for the posed representations and "dpt" otherwise ("linear" is accepted for the
    pose-free ones); adaptor_type overrides the YAML's type with one of its confidence / mask subsets
    (e.g. "pointmap+confidence": reference model.py:407-587 lists all twenty)."""
    if adaptor_config not in ADAPTOR_CONFIGS:
        raise ValueError(f"adaptor_config must be one of {sorted(ADAPTOR_CONFIGS)}, got {adaptor_config!r}")
    ac = copy.deepcopy(ADAPTOR_CONFIGS[adaptor_config])
    dense = ac.pop("dense")
    rep = ac["scene_rep_type"]
    posed = "pose" in rep
    name = ac["type"] + "+scale"
    if adaptor_type is not None:
        if not adaptor_type.startswith(rep):
            raise ValueError(f"adaptor_type {adaptor_type!r} is not a variant of {rep!r}")
        ac["type"] = adaptor_type
        ac["input_dim"] = ac["scene_rep_dim"] + ("confidence" in adaptor_type) + ("mask" in adaptor_type)
    head_type = head_type or ("dpt+pose" if posed else "dpt")
    dense = {"name": name, **dense}
    scale = {"name": name, "mode": "exp", "vmin": 1e-08, "vmax": INF}
    cfg = {"type": head_type, "adaptor_type": ac["type"], "scale_head": {"output_dim": 1}, "scale_adaptor": scale,
           "gradient_checkpointing": False}
    if head_type == "linear":
        cfg["feature_head"] = {"output_dim": ac["input_dim"]}
    else:
        cfg["feature_head"] = {"feature_dim": 256, "hooks": [0, 1, 2, 3], "checkpoint_gradient": False}
        cfg["regressor_head"] = {"output_dim": ac["input_dim"], "checkpoint_gradient": False}
    if posed:
        cfg["pose_head"] = {"num_resconv_block": 2, "rot_representation_dim": 4}
        cfg["dpt_adaptor"] = dense
        cfg["pose_adaptor"] = {"name": name, "cam_trans_mode": "linear", "cam_trans_vmin": -INF, "cam_trans_vmax": INF,
                               "quaternions_mode": "linear", "quaternions_normalize": True, "quaternions_vmin": -INF,
                               "quaternions_vmax": INF}
        ac["dense_pred_init_dict"], ac["pose_pred_init_dict"] = dense, cfg["pose_adaptor"]
    else:
        cfg["adaptor"] = dense
        ac["init_dict"] = dense
    ac["scale_pred_init_dict"] = scale
    cfg["adaptor_config"] = ac   # Hydra nests the chosen adaptor YAML here; model.py:1822 reads it
    return cfg


def tiny_config(img_size: int = 70, enc_dim: int = 128, enc_depth: int = 2, enc_heads: int = 2, info_dim: int = 128,
                info_heads: int = 2, info_depth: int = 4, indices=(1, 2)) -> dict:
    """Same wiring at toy width/depth, for fast CPU tests of the host logic and of oracle-vs-CUDA parity."""
    cfg = mapanything_config()
    cfg["encoder_config"]["vit_kwargs"] = {
        "img_size": img_size, "patch_size": 14, "embed_dim": enc_dim, "depth": enc_depth, "num_heads": enc_heads,
    }
    ma = cfg["info_sharing_config"]["module_args"]
    ma.update({"depth": info_depth, "indices": list(indices), "dim": info_dim, "num_heads": info_heads})
    return cfg
