"""GPU parity of the multi-modal path (BASELINE config 3): input preprocessing kernels, pose /
depth normalisers, the dense + global geometric encoders and their fusion, and infer() end to end with intrinsics + depth +
poses -- against the fp32 CPU oracle on the same weights and seeded inputs."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu

from test_model_gpu import _assert_within, _build, _fmt, _metrics, _rel, _views  # noqa: E402


def _multimodal_views(n, size, seed, drop_depth=(), drop_pose=(), metric=True):
    """SURVEY 8d config C3 at `size` px: pinhole intrinsics, depth_z in [1,4] m, random poses (view 0 = identity)."""
    g = torch.Generator().manual_seed(seed)
    views = _views(n, size, seed)
    for i, v in enumerate(views):
        f = float(torch.empty(1).uniform_(0.8 * size, 1.2 * size, generator=g))
        c = size / 2.0
        v["intrinsics"] = torch.tensor([[[f, 0, c], [0, f, c], [0, 0, 1.0]]])
        if i not in drop_depth:
            v["depth_z"] = torch.empty(1, size, size, 1).uniform_(1.0, 4.0, generator=g)
        if i not in drop_pose:
            q = torch.randn(4, generator=g) * 0.2 + torch.tensor([0.0, 0.0, 0.0, 1.0])
            q = q / q.norm()
            if i == 0:
                q = torch.tensor([0.0, 0.0, 0.0, 1.0])
            x, y, z, w = q.tolist()
            R = torch.tensor([[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
                              [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                              [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]])
            T = torch.eye(4)
            T[:3, :3] = R
            T[:3, 3] = torch.zeros(3) if i == 0 else torch.randn(3, generator=g)
            v["camera_poses"] = T[None]
        v["is_metric_scale"] = torch.tensor([metric])
    return views


def _cuda(views):
    out = []
    for v in views:
        out.append({k: (x.cuda() if torch.is_tensor(x) else x) for k, x in v.items()})
    return out


def test_preprocess_kernels_match_oracle():
    from mapanything_b200.preprocess import preprocess_input_views_for_inference
    from oracle import inference as I

    views = _multimodal_views(3, 70, seed=3)
    # one view with explicit (un-normalised) ray directions instead of intrinsics, one with (quats, trans) tuple poses
    g = torch.Generator().manual_seed(5)
    views[1].pop("intrinsics")
    views[1]["ray_directions"] = torch.randn(1, 70, 70, 3, generator=g) * 0.3 + torch.tensor([0.0, 0.0, 1.5])
    q = torch.tensor([[0.1, -0.2, 0.05, 0.9]])
    views[2]["camera_poses"] = (q / q.norm(), torch.tensor([[0.3, -1.0, 2.0]]))
    ref = I.preprocess_views([dict(v) for v in views])
    got = preprocess_input_views_for_inference(_cuda(views))
    for r, o in zip(ref, got):
        assert set(r.keys()) == set(o.keys()), (sorted(r.keys()), sorted(o.keys()))
        for k in ("ray_directions_cam", "depth_along_ray", "camera_pose_quats", "camera_pose_trans"):
            if k in r:
                err = (o[k].cpu() - r[k]).abs().max().item()
                assert err <= 2e-6 * max(1.0, r[k].abs().max().item()), (k, err)
        assert bool(o["is_metric_scale"].cpu()[0]) == bool(r["is_metric_scale"][0])


def test_pose_inputs_and_depth_factor_match_oracle():
    from mapanything_b200 import ops
    from oracle import geometry as G

    g = torch.Generator().manual_seed(7)
    V = 9
    q = torch.randn(V, 4, generator=g)
    q = q / q.norm(dim=1, keepdim=True)
    t = torch.randn(V, 3, generator=g) * 2
    has = torch.ones(V, dtype=torch.bool)
    has[3] = has[7] = False
    q_rel, t_rel = G.relative_pose_2_to_1(q[:1].expand(V, 4), t[:1].expand(V, 3), q, t)
    ident = torch.tensor([0.0, 0.0, 0.0, 1.0]).expand(V, 4)
    q_ref = torch.where(has[:, None], q_rel, ident)
    t_ref = torch.where(has[:, None], t_rel, torch.zeros(V, 3))
    t_scaled, factor = G.normalize_pose_translations(t_ref[None])
    q8, t8, s8 = ops.pose_inputs(q.cuda(), t.cuda(), has.to(torch.uint8).cuda())
    assert (q8[:, :4].cpu() - q_ref).abs().max().item() < 1e-6
    assert (t8[:, :3].cpu() - t_scaled[0]).abs().max().item() < 1e-5
    assert abs(s8[0, 0].item() - math.log(factor.item() + 1e-8)) < 1e-6
    assert q8[:, 4:].abs().max().item() == 0 and t8[:, 3:].abs().max().item() == 0 and s8[:, 1:].abs().max().item() == 0

    d = torch.empty(3, 70, 70, 1).uniform_(0.5, 6.0, generator=g)
    d[1, :20] = 0.0  # invalid pixels are excluded from the mean
    dn, fac = G.normalize_depth_nonzero(d)
    f, lf8 = ops.depth_factor(d.cuda())
    assert ((f.cpu() - fac).abs() / fac).max().item() < 1e-5
    assert (lf8[:, 0].cpu() - torch.log(fac + 1e-8)).abs().max().item() < 1e-5
    # unshuffle + depth transform: hi + lo reproduces the fp32 value to ~2^-16
    x = ops.unshuffle_split(d.cuda(), 14, 200, f)
    ref = torch.nn.functional.pixel_unshuffle(G.log_of_norm(dn).permute(0, 3, 1, 2), 14).permute(0, 2, 3, 1)  # (3,5,5,196)
    hi, lo, hi2 = x[..., :200].float().cpu(), x[..., 200:400].float().cpu(), x[..., 400:].float().cpu()
    assert torch.equal(hi, hi2)
    assert ((hi + lo)[..., :196] - ref).abs().max().item() < 3e-5
    assert (hi + lo)[..., 196:].abs().max().item() == 0


@pytest.mark.parametrize("init", ["reference", "hard"])
def test_fused_features_with_geometric_inputs(init):
    """Encoder features + ray / depth / pose encoders + fusion LayerNorm vs the oracle's fused features (DPT tap 0)."""
    from mapanything_b200.preprocess import preprocess_input_views_for_inference
    from oracle import inference as I
    from oracle.config import tiny_config

    oracle, model = _build(tiny_config, seed=2, init=init)
    views = _multimodal_views(4, 70, seed=21, drop_depth=(2,), drop_pose=(3,))
    on = {"overall_prob": 1.0, "dropout_prob": 0.0, "ray_dirs_prob": 1.0, "depth_prob": 1.0, "cam_prob": 1.0}
    oracle.geometric_input_config.update(on)
    model.geometric_input_config.update(on)
    with torch.no_grad():
        _, ref = oracle(I.preprocess_views([dict(v) for v in views]), return_internals=True)
        _, ref_img_only = oracle([{"img": v["img"], "data_norm_type": v["data_norm_type"]} for v in views], return_internals=True)
    pv = preprocess_input_views_for_inference(_cuda(views))
    eng = model.engine()
    N = 25
    with torch.no_grad():
        feat = eng.encode(torch.cat([v["img"] for v in pv]))
        model._fuse_geometric_inputs(eng, feat, pv, 0, N, None, None)
        fused = eng.fuse_norm(feat)
    want = ref["fused"].permute(0, 2, 3, 1).reshape(-1, fused.shape[1])
    base = ref_img_only["fused"].permute(0, 2, 3, 1).reshape(-1, fused.shape[1])
    e = _rel(fused, want)
    moved = _rel(base, want)
    print(f"\n[{init}] fused features rel err {e:.3e} (geometric inputs move them by {moved:.3e})")
    assert moved > 5 * e, "the geometric inputs must matter for this test to mean anything"
    assert e < 1e-2


def test_infer_multimodal_matches_oracle():
    """infer() with intrinsics + depth_z + camera_poses on every view (BASELINE config 3 at toy size)."""
    from oracle.config import tiny_config

    oracle, model = _build(tiny_config, seed=3, init="reference")
    views = _multimodal_views(4, 70, seed=31)
    ref = oracle.infer([dict(v) for v in views], apply_mask=False)
    got = model.infer([dict(v) for v in views], apply_mask=False)
    img_only = oracle.infer([{"img": v["img"], "data_norm_type": v["data_norm_type"]} for v in views], apply_mask=False)
    m = _metrics(got, ref)
    print(f"\n[tiny multi-modal infer] geometric inputs move the oracle output by: {_fmt(_metrics(img_only, ref))}")
    _assert_within(m, "tiny, multi-modal infer, V=4")
    for gv, rv in zip(got, ref):
        assert set(gv.keys()) == set(rv.keys())
    # the caller's dicts were moved to the device in place, config restored
    assert model.geometric_input_config["overall_prob"] == 0


def test_infer_ignore_flags_fall_back_to_image_only():
    from oracle.config import tiny_config

    oracle, model = _build(tiny_config, seed=3, init="reference")
    views = _multimodal_views(3, 70, seed=41)
    a = model.infer([dict(v) for v in views], ignore_calibration_inputs=True, ignore_depth_inputs=True, ignore_pose_inputs=True,
                    apply_mask=False)
    b = model.infer([{"img": v["img"], "data_norm_type": v["data_norm_type"]} for v in views], apply_mask=False)
    for x, y in zip(a, b):
        assert torch.equal(x["pts3d"], y["pts3d"])


def test_forward_full_size_multimodal_two_views():
    """BASELINE config 3 at full width (ViT-L, D = 768, 518 px; 2 views to keep the CPU oracle affordable): intrinsics + depth +
    poses through the full-size geometric encoders (588 / 196 -> 588 -> 768 -> 1024 convs, 592-channel padding, split-bf16
    first conv and global MLPs).  Bound: the stated tolerances or 2x the reference's own bf16-autocast floor."""
    from mapanything_b200.preprocess import preprocess_input_views_for_inference
    from oracle import inference as I
    from oracle.config import mapanything_config

    oracle, model = _build(mapanything_config, seed=0, init="reference")
    views = _multimodal_views(2, 518, seed=51)
    on = {"overall_prob": 1.0, "dropout_prob": 0.0, "ray_dirs_prob": 1.0, "depth_prob": 1.0, "cam_prob": 1.0}
    oracle.geometric_input_config.update(on)
    model.geometric_input_config.update(on)
    pv_cpu = I.preprocess_views([dict(v) for v in views])
    with torch.no_grad():
        ref, ref_int = oracle(pv_cpu, return_internals=True)
        amp = oracle(pv_cpu, amp_bf16=True)
        _, img_only = oracle([{"img": v["img"], "data_norm_type": v["data_norm_type"]} for v in views], return_internals=True)
    pv = preprocess_input_views_for_inference(_cuda(views))
    got = model(pv)
    # the geometric inputs must actually move the fused features, and we must follow that move
    eng = model.engine()
    with torch.no_grad():
        feat = eng.encode(torch.cat([v["img"] for v in pv]))
        model._fuse_geometric_inputs(eng, feat, pv, 0, 37 * 37, None, None)
        fused = eng.fuse_norm(feat)
    want = ref_int["fused"].permute(0, 2, 3, 1).reshape(-1, fused.shape[1])
    base = img_only["fused"].permute(0, 2, 3, 1).reshape(-1, fused.shape[1])
    e, moved = _rel(fused, want), _rel(base, want)
    print(f"\n[full-size multi-modal] fused features rel err {e:.3e}; geometric inputs move them by {moved:.3e}")
    assert moved > 5 * e and e < 1.5e-2
    _assert_within(_metrics(got, ref), "full-size multi-modal (V=2), reference-style init", floor=_metrics(amp, ref))


def test_depthmap_to_world_frame_matches_reference_golden():
    """ma_depthmap_to_world vs the reference's own outputs (tests/golden/depthmap.npz): the camera-frame points are
    bit-exact (same operation order, IEEE roundings), the world-frame points within 1e-6 (sum order of the 4-vector product)."""
    from pathlib import Path

    import numpy as np

    from mapanything_b200.geometry import depthmap_to_camera_frame, depthmap_to_world_frame

    gold = np.load(Path(__file__).parent / "golden" / "depthmap.npz")
    depth, K, pose = (torch.from_numpy(gold[k]).cuda() for k in ("depth", "K", "pose"))
    pc, valid = depthmap_to_camera_frame(depth, K)
    assert np.array_equal(pc.cpu().numpy(), gold["pts_cam"])
    assert valid.dtype == torch.bool and np.array_equal(valid.cpu().numpy(), gold["valid"])
    pw, valid2 = depthmap_to_world_frame(depth, K, pose)
    assert np.abs(pw.cpu().numpy() - gold["pts_world"]).max() < 1e-6 and torch.equal(valid, valid2)
    p1, v1 = depthmap_to_world_frame(depth[1], K[1], pose[1])  # un-batched form
    assert p1.shape == (24, 32, 3) and torch.equal(p1, pw[1]) and torch.equal(v1, valid[1])
