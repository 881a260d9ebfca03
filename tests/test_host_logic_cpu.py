"""Host-side logic of the drop-in that needs no GPU: the product's own input validator (same messages and ValueError as the
reference, mapanything/utils/inference.py:128-199), the config helpers that mirror the reference's YAML composition, and the
loud failures of everything that would need a CPU compute path."""
import pytest
import torch


def test_product_validator_error_paths():
    from mapanything_b200.inference import validate_input_views_for_inference as validate

    img = torch.zeros(1, 3, 14, 14)
    base = {"img": img, "data_norm_type": ["dinov2"]}
    assert validate([dict(base)])[0]["img"] is img            # returns the same list / dict objects (reference :199)
    with pytest.raises(ValueError, match="At least one view"):
        validate([])
    with pytest.raises(ValueError, match="invalid keys"):
        validate([{**base, "bogus": 1}])
    with pytest.raises(ValueError, match="missing required"):
        validate([{"img": img}])
    with pytest.raises(ValueError, match="conflicting"):
        validate([{**base, "intrinsics": torch.eye(3)[None], "ray_directions": torch.zeros(1, 14, 14, 3)}])
    with pytest.raises(ValueError, match="depth constraint"):
        validate([{**base, "depth_z": torch.zeros(1, 14, 14, 1)}])
    with pytest.raises(ValueError, match="reference view"):
        validate([dict(base), {**base, "camera_poses": torch.eye(4)[None]}])
    # ranks > 0 of a view-sharded scene: their first local view is not the scene's reference view
    validate([dict(base), {**base, "camera_poses": torch.eye(4)[None]}], first_view_is_reference=False)
    # depth with calibration, pose on view 0: accepted
    ok = [{**base, "intrinsics": torch.eye(3)[None], "depth_z": torch.ones(1, 14, 14, 1), "camera_poses": torch.eye(4)[None]},
          {**base, "ray_directions": torch.zeros(1, 14, 14, 3)}]
    assert validate(ok) is ok


def test_product_validator_matches_oracle_validator_messages():
    """The oracle's validator is pinned to the reference's (tests/test_oracle_golden.py); the product's must say the same."""
    from mapanything_b200.inference import validate_input_views_for_inference as validate
    from oracle import inference as I

    img = torch.zeros(1, 3, 14, 14)
    base = {"img": img, "data_norm_type": ["dinov2"]}
    cases = [[], [{**base, "bogus": 1}], [{"img": img}], [{**base, "depth_z": torch.zeros(1, 14, 14, 1)}],
             [dict(base), {**base, "camera_poses": torch.eye(4)[None]}],
             [{**base, "intrinsics": torch.eye(3)[None], "ray_directions": torch.zeros(1, 14, 14, 3)}]]
    for views in cases:
        with pytest.raises(ValueError) as a:
            validate([dict(v) for v in views])
        with pytest.raises(ValueError) as b:
            I.validate_views([dict(v) for v in views])
        n = min(len(str(a.value)), len(str(b.value)))   # the oracle shortens one message after its first sentence
        assert n >= 30 and str(a.value)[:n] == str(b.value)[:n]


def test_variant_configs_mirror_the_reference_yamls():
    from mapanything_b200.config import ADAPTOR_CONFIGS, INFO_SHARING_VARIANTS, mapanything_variant_config, pred_head_variant_config

    for name, ac in ADAPTOR_CONFIGS.items():
        ph = pred_head_variant_config(name)
        posed = "pose" in ac["scene_rep_type"]
        assert ph["type"] == ("dpt+pose" if posed else "dpt") and ph["adaptor_type"] == ac["type"]
        assert ph["regressor_head"]["output_dim"] == ac["input_dim"] == ac["scene_rep_dim"] + 2
        assert ("dpt_adaptor" in ph and "pose_adaptor" in ph and "pose_head" in ph) == posed
        assert ("adaptor" in ph) == (not posed)
        assert ph["adaptor_config"]["type"] == ac["type"] and ph["scale_adaptor"]["vmin"] == 1e-8
    assert pred_head_variant_config("pointmap_factored_raydirs_depth_pose_confidence_mask_scale")["adaptor_config"][
        "use_factored_predictions_for_global_pointmaps"] is True
    ph = pred_head_variant_config("pointmap_confidence_mask_scale", adaptor_type="pointmap+confidence")
    assert ph["regressor_head"]["output_dim"] == 4 and ph["adaptor_type"] == "pointmap+confidence"
    with pytest.raises(ValueError):
        pred_head_variant_config("pointmap_confidence_mask_scale", adaptor_type="raymap+depth")
    with pytest.raises(ValueError):
        pred_head_variant_config("nonsense")
    cfg = mapanything_variant_config("gat_ifr_24_layers", adaptor_config="campointmap_pose_confidence_mask_scale")
    assert cfg["info_sharing_config"]["model_type"] == "global_attention"
    assert cfg["pred_head_config"]["dpt_adaptor"]["pointmap_mode"] == "z_exp"
    assert set(INFO_SHARING_VARIANTS) >= {"aat_ifr_24_layers", "aat_ifr_48_layers", "gat_ifr_24_layers"}


def test_cpu_module_fails_loudly():
    """No CPU / PyTorch compute path: forward, infer and post-processing raise on a CPU module / CPU tensors."""
    from mapanything_b200 import MapAnything, tiny_config
    from mapanything_b200.inference import postprocess_model_outputs_for_inference

    m = MapAnything(**tiny_config()).eval()
    views = [{"img": torch.zeros(1, 3, 70, 70), "data_norm_type": ["dinov2"]}]
    with pytest.raises(RuntimeError, match="no CPU path"):
        m(views)
    with pytest.raises(RuntimeError, match="no CPU path"):
        m.infer(views)
    with pytest.raises(RuntimeError, match="no CPU path"):
        postprocess_model_outputs_for_inference([{"pts3d": torch.zeros(1, 70, 70, 3)}], views)
    with pytest.raises(RuntimeError):   # parameter containers: an accidental PyTorch forward of a sub-module raises too
        m.scale_head(torch.zeros(1, 128, 1))
