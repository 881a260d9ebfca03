"""The other scene representations / prediction heads of the reference model class (round-1 verdict "missing" 5):
`pred_head_type` linear / dpt / dpt+pose (reference model.py:320-388) and the 20 adaptor types = 5 representations x
{-, confidence} x {-, mask} (model.py:390-588, decode branches :1618-1907).

  * ma_decode_scene against the oracle's adaptor + decode arithmetic on random head outputs: every adaptor type (fp32, tight);
  * whole toy-width models against the oracle for the adaptor YAMLs the reference ships
    (configs/model/pred_head/adaptor_config/*.yaml) and for the linear head;
  * `infer` on a pose-free representation: the keys the reference's post-processing would produce, and only those.
"""
import copy
import itertools

import pytest
import torch

from test_model_gpu import _rel, _rot_err_deg, _views

pytestmark = pytest.mark.gpu

REPS = ["pointmap", "raymap+depth", "raydirs+depth+pose", "campointmap+pose", "pointmap+raydirs+depth+pose"]
ADAPTOR_TYPES = [r + s for r, s in itertools.product(REPS, ["", "+confidence", "+mask", "+confidence+mask"])]


def _oracle_decode(raw_nchw, pose_raw, scale_raw, adaptor_type, cfg, use_factored):
    """The oracle model's decode (oracle/model.py::_decode_other_scene_reps) on given head outputs."""
    from oracle import uniception_modules as U
    from oracle.model import MapAnythingOracle

    class _Shim:
        pass

    shim = _Shim()
    shim.scene_rep_type = adaptor_type
    shim.scene_rep, shim.has_conf, shim.has_mask = U.split_adaptor_type(adaptor_type)
    shim.dense_adaptor_cfg = cfg
    shim.pred_head_config = {"adaptor_config": {"use_factored_predictions_for_global_pointmaps": use_factored}}
    scale = U.scale_adaptor_exp(scale_raw.view(1, 1, 1)).squeeze(-1)
    return MapAnythingOracle._decode_other_scene_reps(shim, raw_nchw, pose_raw, scale, raw_nchw.shape[0], 1)


@pytest.mark.parametrize("adaptor_type", ADAPTOR_TYPES)
@pytest.mark.parametrize("point_mode", ["exp", "z_exp", "linear"])
def test_decode_scene_matches_oracle(adaptor_type, point_mode):
    from mapanything_b200 import ops
    from oracle import uniception_modules as U

    rep, has_conf, has_mask = U.split_adaptor_type(adaptor_type)
    if point_mode != "exp" and rep in ("raymap+depth", "raydirs+depth+pose"):
        pytest.skip("no point channels in this representation")
    n, H, W = 3, 9, 11
    ch = U.SCENE_REP_CHANNELS[rep] + has_conf + has_mask
    g = torch.Generator().manual_seed(sum(map(ord, adaptor_type)))
    raw = torch.randn(n, ch, H, W, generator=g)
    pose_raw = torch.randn(n, 7, generator=g) if "pose" in rep else None
    scale_raw = torch.tensor([0.37])
    for use_factored in ([False, True] if rep == "pointmap+raydirs+depth+pose" else [False]):
        ref = _oracle_decode(raw, pose_raw, scale_raw, adaptor_type, {"pointmap_mode": point_mode, "confidence_vmin": 1},
                             use_factored)
        ld = 8 if ch <= 8 else 12
        raw_rows = torch.zeros(n * H * W, ld)
        raw_rows[:, :ch] = raw.permute(0, 2, 3, 1).reshape(-1, ch)
        got = ops.decode_scene(raw_rows.cuda(), None if pose_raw is None else pose_raw.cuda(), scale_raw.cuda(), n, H, W,
                               rep=rep, has_conf=bool(has_conf), has_mask=bool(has_mask), point_mode=point_mode,
                               use_factored=use_factored)
        assert set(got) == set(ref[0]), (sorted(got), sorted(ref[0]))
        for key in got:
            for i in range(n):
                r = ref[i][key]
                gk = got[key].cpu() if key == "metric_scaling_factor" else got[key][i:i + 1].cpu()
                if r.dtype == torch.bool:
                    assert torch.equal(gk, r), key
                else:
                    torch.testing.assert_close(gk, r, rtol=2e-5, atol=2e-6, msg=lambda m: f"{adaptor_type} {key}: {m}")


def test_decode_scene_released_representation_is_decode_dense():
    """MA_REP_RAYDIRS_DEPTH_POSE with confidence + mask is the released model's ma_decode_dense, bit for bit."""
    from mapanything_b200 import ops

    n, H, W = 2, 14, 14
    g = torch.Generator().manual_seed(5)
    raw = torch.randn(n * H * W, 8, generator=g).cuda()
    pose_raw, scale_raw = torch.randn(n, 7, generator=g).cuda(), torch.tensor([-0.2]).cuda()
    a = ops.decode_dense(raw, pose_raw, scale_raw, n, H, W)
    b = ops.decode_scene(raw, pose_raw, scale_raw, n, H, W, rep="raydirs+depth+pose", has_conf=True, has_mask=True)
    assert set(a) == set(b)
    for k in a:
        assert torch.equal(a[k], b[k]), k


def test_decode_scene_argument_errors():
    from mapanything_b200 import ops
    from mapanything_b200._lib import MapAnythingB200Error

    raw = torch.zeros(4, 8, device="cuda")
    s = torch.zeros(1, device="cuda")
    with pytest.raises(ValueError):
        ops.decode_scene(raw, None, s, 1, 2, 2, rep="nonsense", has_conf=False, has_mask=False)
    with pytest.raises(ValueError):   # posed representation without pose_raw
        ops.decode_scene(raw, None, s, 1, 2, 2, rep="campointmap+pose", has_conf=False, has_mask=False)
    with pytest.raises(MapAnythingB200Error):   # 7 + 2 channels do not fit 8 columns
        ops.decode_scene(raw, None, s, 1, 2, 2, rep="raymap+depth", has_conf=True, has_mask=True)


# ------------------------------------------------------------------------------------------ whole models
def _tiny_with_head(**kw):
    from mapanything_b200.config import pred_head_variant_config
    from oracle.config import tiny_config

    cfg = tiny_config()
    cfg["pred_head_config"] = pred_head_variant_config(**kw)
    return cfg


def _build_cfg(cfg):
    from mapanything_b200 import MapAnything
    from oracle.model import MapAnythingOracle
    from oracle.weights import init_reference_style

    oracle = init_reference_style(MapAnythingOracle(**copy.deepcopy(cfg)).eval(), 0)
    model = MapAnything(**copy.deepcopy(cfg))
    assert set(model.state_dict()) == set(oracle.state_dict())
    model.load_state_dict(oracle.state_dict(), strict=True)
    return oracle, model.cuda().eval()


TOL = {"pts3d": 1e-2, "pts3d_cam": 1e-2, "ray_origins": 1e-2, "ray_directions": 2e-2, "depth_along_ray": 1e-2, "cam_trans": 2e-2,
       "metric_scaling_factor": 1e-2, "conf": 2e-2, "non_ambiguous_mask_logits": 2e-2}


def _errors(got, ref):
    out = {}
    for g, r in zip(got, ref):
        assert set(g) == set(r), (sorted(g), sorted(r))
        for k in r:
            if k == "cam_quats":
                e = _rot_err_deg(g[k], r[k])
            elif r[k].dtype == torch.bool:
                e = (g[k].cpu() != r[k]).float().mean().item()
            else:
                e = _rel(g[k], r[k])
            out[k] = max(out.get(k, 0.0), e)
    return out


def _check_model(cfg, what, n_views=3):
    oracle, model = _build_cfg(cfg)
    views = _views(n_views, 70, seed=47)
    with torch.no_grad():
        ref = oracle([dict(v) for v in views])
        amp = oracle([dict(v) for v in views], amp_bf16=True)
    got = model([{**v, "img": v["img"].cuda()} for v in views])
    ours, floor = _errors(got, ref), _errors(amp, ref)
    print(f"\n[{what}] ours : " + ", ".join(f"{k}={v:.2e}" for k, v in ours.items()))
    print(f"[{what}] AMP-oracle floor: " + ", ".join(f"{k}={v:.2e}" for k, v in floor.items()))
    for k, e in ours.items():
        tol = 0.1 if k == "cam_quats" else 0.02 if k == "non_ambiguous_mask" else TOL[k]
        assert e <= max(tol, 2 * floor[k]), f"{what}: {k} = {e:.4g} exceeds max({tol}, 2 x {floor[k]:.4g})"
    return model, oracle, views


@pytest.mark.parametrize("adaptor_config", [
    "pointmap_confidence_mask_scale", "campointmap_pose_confidence_mask_scale",
    "pointmap_raydirs_depth_pose_confidence_mask_scale", "pointmap_factored_raydirs_depth_pose_confidence_mask_scale",
])
def test_adaptor_yaml_models_match_oracle(adaptor_config):
    _check_model(_tiny_with_head(adaptor_config=adaptor_config), adaptor_config)


def test_dpt_head_pointmap_confidence_only():
    """configs/model/pred_head/dpt.yaml's default adaptor (pointmap_confidence.yaml): no mask channel -> no mask keys."""
    model, _, views = _check_model(_tiny_with_head(adaptor_config="pointmap_confidence_mask_scale",
                                                   adaptor_type="pointmap+confidence"), "dpt + pointmap+confidence")
    out = model([{**v, "img": v["img"].cuda()} for v in views])
    assert set(out[0]) == {"pts3d", "metric_scaling_factor", "conf"}


def test_linear_head_matches_oracle():
    _check_model(_tiny_with_head(adaptor_config="pointmap_confidence_mask_scale", head_type="linear"), "linear head + pointmap")


def test_linear_head_without_intermediate_features():
    """model_return_type "no_intermediate_features" (reference model.py:266-285): the transformer returns its final features
    only, which the linear head consumes; the DPT heads cannot be built on it."""
    from mapanything_b200 import MapAnything

    cfg = _tiny_with_head(adaptor_config="pointmap_confidence_mask_scale", head_type="linear")
    cfg["info_sharing_config"]["model_return_type"] = "no_intermediate_features"
    _check_model(cfg, "linear head, no intermediate features")
    cfg = _tiny_with_head(adaptor_config="pointmap_confidence_mask_scale")
    cfg["info_sharing_config"]["model_return_type"] = "no_intermediate_features"
    with pytest.raises(ValueError):
        MapAnything(**cfg)
    cfg = _tiny_with_head(adaptor_config="pointmap_confidence_mask_scale")
    cfg["info_sharing_config"]["model_type"] = "cross_attention"
    with pytest.raises(ValueError):
        MapAnything(**cfg)


def test_raymap_depth_model_matches_oracle():
    """No YAML of the reference selects raymap+depth; the model class accepts it (model.py:423-441)."""
    from oracle.config import tiny_config

    cfg = tiny_config()
    ph = cfg["pred_head_config"]
    ph["type"] = "dpt"
    ph["adaptor_type"] = "raymap+depth+confidence"
    ph["regressor_head"]["output_dim"] = 8
    ph["adaptor"] = {"name": "raymap+depth+confidence", "confidence_type": "exp", "confidence_vmin": 1}
    _check_model(cfg, "dpt + raymap+depth+confidence")


def test_posed_representation_needs_the_pose_head():
    from mapanything_b200 import MapAnything

    cfg = _tiny_with_head(adaptor_config="campointmap_pose_confidence_mask_scale", head_type="dpt")
    with pytest.raises((AssertionError, KeyError)):
        MapAnything(**cfg)
    cfg = _tiny_with_head(adaptor_config="pointmap_confidence_mask_scale")
    cfg["pred_head_config"]["regressor_head"]["output_dim"] = 6
    with pytest.raises(ValueError):   # 5 channels consumed, 6 produced
        MapAnything(**cfg)
    cfg = _tiny_with_head(adaptor_config="pointmap_confidence_mask_scale")
    cfg["pred_head_config"]["type"] = "mlp"
    with pytest.raises(ValueError):
        MapAnything(**cfg)


def test_infer_pose_free_representation():
    """infer on a pointmap model: no depth_z / intrinsics / camera_poses; the non-ambiguous mask still zeroes pts3d; the
    edge mask needs depth_z, which this representation does not have -- the reference raises KeyError there
    (mapanything/utils/inference.py:440), so does this."""
    model, oracle, views = _check_model(_tiny_with_head(adaptor_config="pointmap_confidence_mask_scale"), "pointmap for infer")
    with pytest.raises(KeyError):
        model.infer([dict(v) for v in views])
    got = model.infer([dict(v) for v in views], mask_edges=False)
    ref = oracle.infer([dict(v) for v in views], mask_edges=False)
    assert set(got[0]) == set(ref[0]) == {"pts3d", "metric_scaling_factor", "conf", "non_ambiguous_mask",
                                          "non_ambiguous_mask_logits", "img_no_norm", "mask"}
    for g, r in zip(got, ref):
        m = g["mask"].cpu()
        assert (m != r["mask"]).float().mean() < 0.02
        both = (m & r["mask"]).expand_as(r["pts3d"])
        assert torch.all(g["pts3d"].cpu()[~m.expand_as(r["pts3d"])] == 0)
        d = (g["pts3d"].cpu() - r["pts3d"])[both].abs().max() / r["pts3d"][both].abs().max()
        assert d < 2e-2, d
        torch.testing.assert_close(g["img_no_norm"].cpu(), r["img_no_norm"], rtol=0, atol=1e-6)
