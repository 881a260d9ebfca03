"""CPU tests: the oracle restatement vs golden vectors produced by the REFERENCE's own code (oracle/make_golden.py).

These pin the oracle for everything the reference ships source for: the DINOv2 ViT, the geometry math and the
infer() pre/post-processing.  (The uniception modules have no pin -- see oracle/__init__.py.)
"""
from pathlib import Path

import numpy as np
import pytest
import torch

GOLD = Path(__file__).parent / "golden"


def _t(a):
    return torch.from_numpy(np.asarray(a))


@pytest.fixture(scope="module")
def geo():
    return np.load(GOLD / "geometry.npz")


def test_rays_and_intrinsics(geo):
    from oracle import geometry as G

    rays = G.rays_from_intrinsics(_t(geo["K"]), 56, 70)
    assert torch.equal(rays, _t(geo["rays"]))
    assert torch.allclose(G.intrinsics_from_rays(_t(geo["rays"])), _t(geo["K_rec"]), atol=1e-4)
    # round trip recovers the true intrinsics (f=500 -> 499.9997)
    assert torch.allclose(_t(geo["K_rec"]), _t(geo["K"]), atol=5e-2)
    big = G.rays_from_intrinsics(torch.tensor([[[900.0, 0, 640.0], [0, 905.0, 400.0], [0, 0, 1]]]), 800, 1300)
    assert torch.allclose(G.intrinsics_from_rays(big), _t(geo["K_rec_big"]), atol=1e-3)


def test_quaternions(geo):
    from oracle import geometry as G

    q, q2, t1, t2 = (_t(geo[k]) for k in ("q", "q2", "t1", "t2"))
    assert torch.equal(G.quat_to_rotmat(q), _t(geo["rotmat"]))
    assert torch.equal(G.rotmat_to_quat(_t(geo["rotmat"])), _t(geo["q_back"]))
    assert torch.equal(G.quat_inverse(q), _t(geo["q_inv"]))
    assert torch.equal(G.quat_multiply(q, q2), _t(geo["q_mul"]))
    qn, q2n = q / q.norm(dim=1, keepdim=True), q2 / q2.norm(dim=1, keepdim=True)
    rq, rt = G.relative_pose_2_to_1(qn, t1, q2n, t2)
    assert torch.allclose(rq, _t(geo["rel_q"]), atol=1e-6) and torch.allclose(rt, _t(geo["rel_t"]), atol=1e-6)
    # known answers: identity, and q (x) q^-1 = identity
    eye = G.quat_to_rotmat(torch.tensor([0.0, 0.0, 0.0, 1.0]))
    assert torch.equal(eye, torch.eye(3))
    ident = G.quat_multiply(qn, G.quat_inverse(qn))
    assert torch.allclose(ident, torch.tensor([0.0, 0.0, 0.0, 1.0]).expand_as(ident), atol=1e-6)
    # rotmat -> quat -> rotmat round trip, w >= 0
    qb = G.rotmat_to_quat(G.quat_to_rotmat(q))
    assert (qb[:, 3] >= 0).all()
    assert torch.allclose(G.quat_to_rotmat(qb), G.quat_to_rotmat(q), atol=1e-5)


def test_pointmap_and_normalisers(geo):
    from oracle import geometry as G

    pts = G.pointmap_from_rays_depth_pose(_t(geo["rays"]), _t(geo["depth"]), _t(geo["t1"])[:2], _t(geo["q"])[:2])
    assert torch.allclose(pts, _t(geo["pts_world"]), atol=1e-6)
    dn, df = G.normalize_depth_nonzero(_t(geo["depth_sparse"]))
    assert torch.equal(dn, _t(geo["depth_norm"])) and torch.equal(df, _t(geo["depth_factor"]))
    tn, tf = G.normalize_pose_translations(_t(geo["trans_views"]))
    assert torch.equal(tn, _t(geo["trans_norm"])) and torch.equal(tf, _t(geo["trans_factor"]))
    assert torch.equal(G.log_of_norm(_t(geo["depth_sparse"])), _t(geo["depth_log"]))
    # empty (all-zero) depth: factor clips to 1e-8, output stays finite
    z = torch.zeros(1, 4, 4, 1)
    zn, zf = G.normalize_depth_nonzero(z)
    assert torch.isfinite(zn).all() and zf.item() == pytest.approx(1e-8)


def test_edge_masks_bit_exact(geo):
    from oracle import geometry as G

    n, nm = G.points_to_normals(geo["edge_pts"], geo["edge_mask"])
    assert np.array_equal(nm, geo["edge_nmask"])
    assert np.allclose(n, geo["edge_normals"], atol=1e-7)
    assert np.array_equal(G.normals_edge(geo["edge_normals"], 5.0, geo["edge_nmask"]), geo["edge_ne"])
    dz = geo["edge_pts"][..., 2]
    assert np.array_equal(G.depth_edge(dz, 0.03, geo["edge_mask"]), geo["edge_de"])
    assert geo["edge_ne"].sum() > 0 and geo["edge_de"].sum() > 0  # the fixture actually exercises both


def test_preprocess_matches_reference():
    from oracle import inference as I

    g = np.load(GOLD / "inference.npz")
    img, k, dz, pose, q, rays = (_t(g[n]) for n in ("pre_img", "pre_K", "pre_depth_z", "pre_pose", "pre_q", "pre_rays_in"))
    views = [
        {"img": img, "data_norm_type": ["dinov2"], "intrinsics": k, "depth_z": dz, "camera_poses": pose},
        {"img": img, "data_norm_type": ["dinov2"], "ray_directions": rays,
         "camera_poses": (q, torch.tensor([[0.5, 0.5, 0.5]])), "is_metric_scale": torch.tensor([False])},
    ]
    out = I.preprocess_views(I.validate_views(views))
    for i in range(2):
        for key in ("ray_directions_cam", "depth_along_ray", "camera_pose_quats", "camera_pose_trans", "is_metric_scale"):
            name = f"pre{i}_{key}"
            if name in g.files:
                assert torch.allclose(out[i][key].float(), _t(g[name]).float(), atol=1e-6), name
    assert "depth_along_ray" not in out[1] and "intrinsics" not in out[0] and "camera_poses" not in out[0]


@pytest.mark.parametrize("tag,kw", [("default", {}), ("conf", {"apply_confidence_mask": True, "confidence_percentile": 25}),
                                    ("noedge", {"mask_edges": False})])
def test_postprocess_matches_reference(tag, kw):
    from oracle import inference as I

    g = np.load(GOLD / "inference.npz")
    raw = {k[4:]: _t(g[k]) for k in g.files if k.startswith("raw_") and k != "raw_img"}
    out = I.postprocess_outputs([raw], [{"img": _t(g["raw_img"]), "data_norm_type": ["dinov2"]}], **kw)[0]
    keys = [k[len(f"post_{tag}_"):] for k in g.files if k.startswith(f"post_{tag}_")]
    assert set(keys) == set(out.keys())
    for key in keys:
        ref = _t(g[f"post_{tag}_{key}"])
        if ref.dtype == torch.bool:
            assert torch.equal(out[key], ref), key
        else:
            assert torch.allclose(out[key], ref, atol=1e-3 if key == "intrinsics" else 1e-6), key
    m = out["mask"]
    assert 0 < m.sum() < m.numel()


def test_validate_errors():
    from oracle import inference as I

    img = torch.zeros(1, 3, 14, 14)
    base = {"img": img, "data_norm_type": ["dinov2"]}
    with pytest.raises(ValueError, match="At least one view"):
        I.validate_views([])
    with pytest.raises(ValueError, match="invalid keys"):
        I.validate_views([{**base, "bogus": 1}])
    with pytest.raises(ValueError, match="missing required"):
        I.validate_views([{"img": img}])
    with pytest.raises(ValueError, match="conflicting"):
        I.validate_views([{**base, "intrinsics": torch.eye(3)[None], "ray_directions": torch.zeros(1, 14, 14, 3)}])
    with pytest.raises(ValueError, match="depth constraint"):
        I.validate_views([{**base, "depth_z": torch.zeros(1, 14, 14, 1)}])
    with pytest.raises(ValueError, match="reference view"):
        I.validate_views([base, {**base, "camera_poses": torch.eye(4)[None]}])


def test_small_vit_matches_reference_square_and_rect():
    from oracle.vit import OracleDinoV2
    from oracle.weights import synth_state_dict

    g = np.load(GOLD / "vit.npz")
    m = OracleDinoV2(img_size=70, embed_dim=128, depth=2, num_heads=2).eval()
    m.load_state_dict(synth_state_dict(m, seed=3))
    with torch.no_grad():
        for tag in ("sq", "rect"):
            out = m.forward_patch_tokens(_t(g[f"small_{tag}_in"]))
            assert torch.allclose(out, _t(g[f"small_{tag}_out"]), atol=2e-5), tag


def test_vit_large_matches_reference_518():
    """Full ViT-L/14 at 518 px on one view (~5 s on 8 cores): checks sampled tokens against the reference run."""
    from oracle.vit import OracleDinoV2
    from oracle.weights import synth_state_dict

    g = np.load(GOLD / "vit.npz")
    m = OracleDinoV2().eval()
    m.load_state_dict(synth_state_dict(m, seed=0))
    gen = torch.Generator().manual_seed(1234)
    img = torch.randn(1, 3, 518, 518, generator=gen)
    with torch.no_grad():
        out = m.forward_patch_tokens(img)
    rows = g["vitl_rows"].tolist()
    assert torch.allclose(out[0, rows], _t(g["vitl_tokens"]), atol=5e-4)
    assert torch.allclose(out[0].mean(0), _t(g["vitl_col_mean"]), atol=5e-4)
    assert out.abs().mean().item() == pytest.approx(float(g["vitl_mean_abs"]), rel=1e-3)


def test_depthmap_to_world_frame_matches_reference():
    from oracle import geometry as G

    gold = np.load(GOLD / "depthmap.npz")
    depth, K, pose = (torch.from_numpy(gold[k]) for k in ("depth", "K", "pose"))
    pc, valid = G.depthmap_to_camera_frame(depth, K)
    assert np.array_equal(pc.numpy(), gold["pts_cam"]) and np.array_equal(valid.numpy(), gold["valid"])
    pw, _ = G.depthmap_to_world_frame(depth, K, pose)
    assert np.abs(pw.numpy() - gold["pts_world"]).max() < 1e-6


def test_small_vit_matches_huggingface_dinov2():
    """Independent cross-check of stage 1: the oracle ViT against `transformers.Dinov2Model` -- a second
    implementation of DINOv2 that shares no code with the reference's vendored copy -- on the same synthetic weights."""
    transformers = pytest.importorskip("transformers")
    from oracle.vit import OracleDinoV2
    from oracle.weights import synth_state_dict

    dim, depth, heads = 128, 2, 2
    m = OracleDinoV2(img_size=70, embed_dim=dim, depth=depth, num_heads=heads).eval()
    sd = synth_state_dict(m, seed=5)
    m.load_state_dict(sd)
    cfg = transformers.Dinov2Config(hidden_size=dim, num_hidden_layers=depth, num_attention_heads=heads, image_size=70,
                                    patch_size=14, mlp_ratio=4, qkv_bias=True, hidden_act="gelu", layer_norm_eps=1e-6,
                                    layerscale_value=1.0, attn_implementation="eager")
    hf = transformers.Dinov2Model(cfg).eval()
    new = {"embeddings.cls_token": sd["cls_token"], "embeddings.mask_token": sd["mask_token"],
           "embeddings.position_embeddings": sd["pos_embed"],
           "embeddings.patch_embeddings.projection.weight": sd["patch_embed.proj.weight"],
           "embeddings.patch_embeddings.projection.bias": sd["patch_embed.proj.bias"],
           "layernorm.weight": sd["norm.weight"], "layernorm.bias": sd["norm.bias"]}
    for i in range(depth):
        s, d = f"blocks.{i}.", f"encoder.layer.{i}."
        qw, kw, vw = sd[s + "attn.qkv.weight"].chunk(3, 0)
        qb, kb, vb = sd[s + "attn.qkv.bias"].chunk(3, 0)
        for n, w, b in (("query", qw, qb), ("key", kw, kb), ("value", vw, vb)):
            new[d + f"attention.attention.{n}.weight"], new[d + f"attention.attention.{n}.bias"] = w, b
        new[d + "attention.output.dense.weight"], new[d + "attention.output.dense.bias"] = sd[s + "attn.proj.weight"], sd[s + "attn.proj.bias"]
        new[d + "layer_scale1.lambda1"], new[d + "layer_scale2.lambda1"] = sd[s + "ls1.gamma"], sd[s + "ls2.gamma"]
        for n in ("norm1", "norm2", "mlp.fc1", "mlp.fc2"):
            new[d + n + ".weight"], new[d + n + ".bias"] = sd[s + n + ".weight"], sd[s + n + ".bias"]
    missing, unexpected = hf.load_state_dict(new, strict=False)
    assert not unexpected and not [k for k in missing if "mask_token" not in k], (missing, unexpected)
    img = torch.randn(2, 3, 70, 70, generator=torch.Generator().manual_seed(9))
    with torch.no_grad():
        ours = m.forward_patch_tokens(img)
        theirs = hf(pixel_values=img).last_hidden_state[:, 1:]
    assert torch.allclose(ours, theirs, atol=2e-5), (ours - theirs).abs().max()


def test_adaptor_known_answers():
    """Known-answer tests for the restated adaptor arithmetic of the other scene representations (the published DUSt3R /
    MoGe forms the uniception adaptors wrap; oracle/uniception_modules.py): hand-computed values."""
    import math

    from oracle import uniception_modules as U

    assert U.split_adaptor_type("pointmap+raydirs+depth+pose+confidence+mask") == ("pointmap+raydirs+depth+pose", True, True)
    assert U.split_adaptor_type("raymap+depth+mask") == ("raymap+depth", False, True)
    assert U.split_adaptor_type("campointmap+pose") == ("campointmap+pose", False, False)
    with pytest.raises(ValueError):
        U.split_adaptor_type("pointcloud+confidence")
    x = torch.tensor([3.0, 0.0, 4.0]).view(1, 3, 1, 1)                       # norm 5
    p = U.point_activation(x, "exp").flatten()
    assert torch.allclose(p, torch.tensor([0.6, 0.0, 0.8]) * math.expm1(5.0), rtol=1e-6)
    p = U.point_activation(torch.tensor([0.5, -2.0, math.log(3.0)]).view(1, 3, 1, 1), "z_exp").flatten()
    assert torch.allclose(p, torch.tensor([1.5, -6.0, 3.0]), rtol=1e-6)
    assert torch.equal(U.point_activation(x, "linear"), x)
    # pointmap+confidence+mask: channels [xyz | confidence logit | mask logit]
    raw = torch.tensor([0.0, 0.0, 2.0, math.log(2.0), -1.0]).view(1, 5, 1, 1)
    value, conf, mask, logits = U.dense_adaptor(raw, "pointmap+confidence+mask", {"pointmap_mode": "exp", "confidence_vmin": 1})
    assert torch.allclose(value.flatten(), torch.tensor([0.0, 0.0, math.expm1(2.0)]), rtol=1e-6)
    assert abs(conf.item() - 3.0) < 1e-6 and abs(mask.item() - 1 / (1 + math.e)) < 1e-6 and logits.item() == -1.0
    # raymap+depth: origin linear, direction normalised, depth = exp
    raw = torch.tensor([1.0, 2.0, 3.0, 0.0, 3.0, 4.0, math.log(2.0)]).view(1, 7, 1, 1)
    value, conf, mask, logits = U.dense_adaptor(raw, "raymap+depth", {})
    assert conf is None and mask is None and logits is None
    assert torch.allclose(value.flatten(), torch.tensor([1.0, 2.0, 3.0, 0.0, 0.6, 0.8, 2.0]), rtol=1e-6)
    # the released adaptor through the generic entry = the dedicated function
    g = torch.Generator().manual_seed(0)
    raw = torch.randn(2, 6, 3, 3, generator=g)
    a = U.dense_adaptor(raw, "raydirs+depth+pose+confidence+mask", {"confidence_vmin": 1})
    b = U.dense_adaptor_raydirs_depth_conf_mask(raw)
    assert all(torch.equal(x, y) for x, y in zip(a, b))
    # linear head: pixel shuffle of a 1x1 conv
    head = U.LinearFeature(input_feature_dim=4, output_dim=2, patch_size=3)
    out = head(torch.randn(1, 4, 2, 2, generator=g))
    assert out.shape == (1, 2, 6, 6)
