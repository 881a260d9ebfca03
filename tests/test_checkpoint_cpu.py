"""CPU tests of the checkpoint surface the reference exposes (`MapAnything.from_pretrained`, HF layout config.json +
model.safetensors: reference scripts/gradio_app.py:70-71, mapanything/utils/hf_utils/hf_helpers.py:139-154) and of
tools/verify_checkpoint.py, which reads a checkpoint's tensor names / shapes and reports which App. A assumptions they decide."""
import subprocess
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]


def _tiny_model():
    from mapanything_b200 import MapAnything, tiny_config

    torch.manual_seed(0)
    return MapAnything(**tiny_config())


def test_save_pretrained_from_pretrained_round_trip(tmp_path):
    from mapanything_b200 import MapAnything

    m = _tiny_model()
    m.save_pretrained(tmp_path)
    assert (tmp_path / "config.json").exists() and (tmp_path / "model.safetensors").exists()
    m2 = MapAnything.from_pretrained(str(tmp_path))
    a, b = m.state_dict(), m2.state_dict()
    assert set(a) == set(b) and len(a) == 318
    assert all(torch.equal(a[k], b[k]) for k in a)
    # the constructor arguments travelled through config.json (the config dicts the ctor mutates are stored post-mutation)
    assert m2.info_sharing.dim == m.info_sharing.dim and m2.encoder.enc_embed_dim == m.encoder.enc_embed_dim
    assert m2.geometric_input_config["ray_dirs_encoder_config"]["patch_size"] == 14


def test_pth_checkpoint_path_ctor_argument(tmp_path):
    """reference model.py:590-620: pretrained_checkpoint_path = a torch.save'd dict with a "model" entry."""
    from mapanything_b200 import MapAnything, tiny_config

    m = _tiny_model()
    torch.save({"model": m.state_dict()}, tmp_path / "ckpt.pth")
    m2 = MapAnything(**tiny_config(), pretrained_checkpoint_path=str(tmp_path / "ckpt.pth"))
    assert all(torch.equal(v, m2.state_dict()[k]) for k, v in m.state_dict().items())


def test_verify_checkpoint_tool(tmp_path):
    m = _tiny_model()
    m.save_pretrained(tmp_path)
    r = subprocess.run([sys.executable, str(ROOT / "tools" / "verify_checkpoint.py"), str(tmp_path)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "strict load OK" in r.stdout and "info.dim" in r.stdout
    # without config.json the configuration is inferred from the tensors alone
    (tmp_path / "config.json").unlink()
    r = subprocess.run([sys.executable, str(ROOT / "tools" / "verify_checkpoint.py"), str(tmp_path)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    # a checkpoint with a different regressor width is detected and still loads through the inferred config
    from mapanything_b200 import MapAnything, tiny_config

    cfg = tiny_config()
    cfg["pred_head_config"]["regressor_head"]["hidden_dims"] = [128, 32]
    torch.save({"model": MapAnything(**cfg).state_dict()}, tmp_path / "other.pth")
    r = subprocess.run([sys.executable, str(ROOT / "tools" / "verify_checkpoint.py"), str(tmp_path / "other.pth")],
                       capture_output=True, text=True)
    assert r.returncode == 0 and "DIFFERS" in r.stdout and "[128, 32]" in r.stdout, r.stdout[-2000:]


def _tiny_variant(**kw):
    from mapanything_b200 import MapAnything, pred_head_variant_config, tiny_config

    cfg = tiny_config()
    cfg["pred_head_config"] = pred_head_variant_config(**kw)
    torch.manual_seed(0)
    return MapAnything(**cfg), cfg


def test_other_heads_and_scene_representations_construct_on_cpu():
    """pred_head_type linear / dpt / dpt+pose and the adaptor YAMLs of the reference (model.py:320-588): module trees, key
    prefixes and the config mutations the reference constructor performs; invalid combinations raise like the reference."""
    import pytest

    from mapanything_b200 import MapAnything, pred_head_variant_config, tiny_config
    from mapanything_b200.config import ADAPTOR_CONFIGS

    for name, ac in ADAPTOR_CONFIGS.items():
        m, cfg = _tiny_variant(adaptor_config=name)
        keys = set(m.state_dict())
        posed = "pose" in ac["scene_rep_type"]
        assert any(k.startswith("pose_head.") for k in keys) == posed, name
        assert m.pred_head_type == ("dpt+pose" if posed else "dpt") and m.scene_rep_type == ac["type"]
        assert m.state_dict()["dpt_regressor_head.conv2.2.weight"].shape[0] == ac["input_dim"]
        assert any(k.startswith("dense_head.1.") for k in keys) and any(k.startswith("scale_head.") for k in keys)
        # the constructor wrote the dependent sizes into the caller's dict (reference model.py:338-352)
        assert m.pred_head_config["regressor_head"]["input_feature_dim"] == 256
        assert m.pred_head_config["scale_head"]["input_feature_dim"] == m.info_sharing.dim
    m, _ = _tiny_variant(adaptor_config="pointmap_confidence_mask_scale", head_type="linear")
    assert m.state_dict()["dense_head.proj.weight"].shape == (5 * 14 * 14, m.info_sharing.dim, 1, 1)
    assert not any(k.startswith(("dpt_", "pose_head.")) for k in m.state_dict())
    assert m.pred_head_config["feature_head"]["input_feature_dim"] == m.info_sharing.dim
    # all twenty adaptor types parse; the channel count must match the head
    for rep, ch in (("pointmap", 3), ("raymap+depth", 7)):
        for suffix, extra in (("", 0), ("+confidence", 1), ("+mask", 1), ("+confidence+mask", 2)):
            cfg = tiny_config()
            cfg["pred_head_config"].update({"type": "dpt", "adaptor_type": rep + suffix, "adaptor": {"name": rep + suffix}})
            cfg["pred_head_config"]["regressor_head"]["output_dim"] = ch + extra
            assert MapAnything(**cfg).scene_rep_type == rep + suffix
    cfg = tiny_config()
    cfg["pred_head_config"]["adaptor_type"] = "voxels"
    with pytest.raises(ValueError, match="Invalid adaptor_type"):
        MapAnything(**cfg)
    cfg = tiny_config()
    cfg["pred_head_config"] = pred_head_variant_config("campointmap_pose_confidence_mask_scale", head_type="dpt")
    with pytest.raises(AssertionError, match="dpt \\+ pose head"):
        MapAnything(**cfg)
    cfg = tiny_config()
    cfg["pred_head_config"]["dpt_adaptor"]["depth_mode"] = "square"   # not a value the fused decode implements
    with pytest.raises(ValueError, match="fused decode"):
        MapAnything(**cfg)


def test_verify_checkpoint_tool_other_heads(tmp_path):
    """Without config.json the tool infers the head type and an adaptor type that fits the channel count."""
    for i, kw in enumerate((dict(adaptor_config="pointmap_confidence_mask_scale", head_type="linear"),
                            dict(adaptor_config="pointmap_confidence_mask_scale"),
                            dict(adaptor_config="pointmap_raydirs_depth_pose_confidence_mask_scale"))):
        m, _ = _tiny_variant(**kw)
        path = tmp_path / f"v{i}.pth"
        torch.save({"model": m.state_dict()}, path)
        r = subprocess.run([sys.executable, str(ROOT / "tools" / "verify_checkpoint.py"), str(path)], capture_output=True, text=True)
        assert r.returncode == 0 and "strict load OK" in r.stdout, r.stdout[-3000:] + r.stderr[-2000:]
        assert f"head.type                          {m.pred_head_type}" in r.stdout, r.stdout[-3000:]
