"""Generates tests/golden/*.npz by running THE REFERENCE'S OWN CODE, imported read-only from /root/reference.

Run in the build container only (the GPU box has no /root/reference):

    python -m oracle.make_golden

What is pinned: the vendored DINOv2 ViT, every geometry function on the inference path and
the infer() pre/post-processing.  The reference's package __init__ files need omegaconf / uniception, which are
not installed, so `mapanything` and `mapanything.models` are registered as bare namespace modules and the one
uniception symbol the utils import (IMAGE_NORMALIZATION_DICT) is stubbed with the ImageNet statistics.
While generating, the oracle restatement is checked against the reference on the same inputs; the script
fails if they disagree.
"""
from __future__ import annotations

import sys
import types
from pathlib import Path

import numpy as np
import torch

REF = Path("/root/reference")
OUT = Path(__file__).resolve().parents[1] / "tests" / "golden"


def _install_reference_stubs():
    for name, path in [("mapanything", REF / "mapanything"), ("mapanything.models", REF / "mapanything" / "models")]:
        m = types.ModuleType(name)
        m.__path__ = [str(path)]
        sys.modules[name] = m
    for name in ["uniception", "uniception.models", "uniception.models.encoders",
                 "uniception.models.encoders.image_normalizations"]:
        sys.modules.setdefault(name, types.ModuleType(name))

    class _Norm:
        def __init__(self, mean, std):
            self.mean, self.std = torch.tensor(mean), torch.tensor(std)

    sys.modules["uniception.models.encoders.image_normalizations"].IMAGE_NORMALIZATION_DICT = {
        "dinov2": _Norm((0.485, 0.456, 0.406), (0.229, 0.224, 0.225)),
    }


def _assert_close(a, b, tol, what):
    a, b = torch.as_tensor(a).double(), torch.as_tensor(b).double()
    err = (a - b).abs().max().item() if a.numel() else 0.0
    print(f"  oracle vs reference  {what:<44s} max|diff| = {err:.3e}")
    assert err <= tol, f"{what}: oracle deviates from the reference by {err}"


def golden_vit():
    from mapanything.models.external.dinov2.hub.backbones import dinov2_vitl14
    from mapanything.models.external.dinov2.models import vision_transformer as ref_vits

    from oracle.vit import OracleDinoV2
    from oracle.weights import synth_state_dict

    out = {}
    # (a) full ViT-L/14 @ 518, one view
    ref = dinov2_vitl14(pretrained=False).eval()
    sd = synth_state_dict(ref, seed=0)
    ref.load_state_dict(sd)
    mine = OracleDinoV2().eval()
    mine.load_state_dict(sd)
    g = torch.Generator().manual_seed(1234)
    img = torch.randn(1, 3, 518, 518, generator=g)
    with torch.no_grad():
        r = ref.forward_features(img)["x_norm_patchtokens"]
        o = mine.forward_patch_tokens(img)
    _assert_close(o, r, 2e-4, "ViT-L/14 518px x_norm_patchtokens")
    rows = [0, 1, 36, 37, 684, 1000, 1368]
    out["vitl_rows"] = np.array(rows)
    out["vitl_tokens"] = r[0, rows].numpy()
    out["vitl_mean_abs"] = np.array(r.abs().mean().item())
    out["vitl_col_mean"] = r[0].mean(0).numpy()
    del ref, mine

    # (b) small ViT, square (no interpolation) and non-square (bicubic pos-embed path), full outputs
    kw = dict(img_size=70, patch_size=14, embed_dim=128, depth=2, num_heads=2, mlp_ratio=4, init_values=1.0,
              ffn_layer="mlp", block_chunks=0)
    ref = ref_vits.DinoVisionTransformer(**kw).eval()
    sd = synth_state_dict(ref, seed=3)
    ref.load_state_dict(sd)
    mine = OracleDinoV2(img_size=70, embed_dim=128, depth=2, num_heads=2).eval()
    mine.load_state_dict(sd)
    for tag, shape in [("sq", (2, 3, 70, 70)), ("rect", (1, 3, 70, 98))]:
        x = torch.randn(*shape, generator=g)
        with torch.no_grad():
            r = ref.forward_features(x)["x_norm_patchtokens"]
            o = mine.forward_patch_tokens(x)
        _assert_close(o, r, 1e-4, f"small ViT {tag}")
        out[f"small_{tag}_in"] = x.numpy()
        out[f"small_{tag}_out"] = r.numpy()
    np.savez_compressed(OUT / "vit.npz", **out)


def golden_geometry():
    import mapanything.utils.geometry as RG

    import oracle.geometry as G

    g = torch.Generator().manual_seed(7)
    out = {}
    b, h, w = 2, 56, 70
    k = torch.tensor([[[61.0, 0, 34.2], [0, 59.5, 27.1], [0, 0, 1]], [[80.0, 0, 35.0], [0, 80.0, 28.0], [0, 0, 1]]])
    _, rays = RG.get_rays_in_camera_frame(k, h, w, normalize_to_unit_sphere=True)
    _assert_close(G.rays_from_intrinsics(k, h, w), rays, 1e-6, "get_rays_in_camera_frame")
    out["K"], out["rays"] = k.numpy(), rays.numpy()
    kr = RG.recover_pinhole_intrinsics_from_ray_directions(rays)
    _assert_close(G.intrinsics_from_rays(rays), kr, 1e-3, "recover_pinhole_intrinsics (lstsq)")
    out["K_rec"] = kr.numpy()
    # > 1 MP path (five key pixels)
    _, rays_big = RG.get_rays_in_camera_frame(torch.tensor([[[900.0, 0, 640.0], [0, 905.0, 400.0], [0, 0, 1]]]), 800, 1300, True)
    kb = RG.recover_pinhole_intrinsics_from_ray_directions(rays_big)
    _assert_close(G.intrinsics_from_rays(rays_big), kb, 1e-3, "recover_pinhole_intrinsics (>1MP)")
    out["K_rec_big"] = kb.numpy()

    q = torch.randn(6, 4, generator=g)
    q2 = torch.randn(6, 4, generator=g)
    t1, t2 = torch.randn(6, 3, generator=g), torch.randn(6, 3, generator=g)
    out["q"], out["q2"], out["t1"], out["t2"] = q.numpy(), q2.numpy(), t1.numpy(), t2.numpy()
    rm = RG.quaternion_to_rotation_matrix(q)
    _assert_close(G.quat_to_rotmat(q), rm, 1e-6, "quaternion_to_rotation_matrix")
    out["rotmat"] = rm.numpy()
    qb = RG.rotation_matrix_to_quaternion(rm)
    _assert_close(G.rotmat_to_quat(rm), qb, 1e-6, "rotation_matrix_to_quaternion")
    out["q_back"] = qb.numpy()
    _assert_close(G.quat_inverse(q), RG.quaternion_inverse(q), 1e-6, "quaternion_inverse")
    out["q_inv"] = RG.quaternion_inverse(q).numpy()
    _assert_close(G.quat_multiply(q, q2), RG.quaternion_multiply(q, q2), 1e-6, "quaternion_multiply")
    out["q_mul"] = RG.quaternion_multiply(q, q2).numpy()
    qn, q2n = q / q.norm(dim=1, keepdim=True), q2 / q2.norm(dim=1, keepdim=True)
    qr, tr = RG.transform_pose_using_quats_and_trans_2_to_1(qn, t1, q2n, t2)
    mq, mt = G.relative_pose_2_to_1(qn, t1, q2n, t2)
    _assert_close(mq, qr, 1e-5, "transform_pose_2_to_1 quats")
    _assert_close(mt, tr, 1e-5, "transform_pose_2_to_1 trans")
    out["rel_q"], out["rel_t"] = qr.numpy(), tr.numpy()

    depth = torch.rand(b, h, w, 1, generator=g) * 3 + 0.5
    pts = RG.convert_ray_dirs_depth_along_ray_pose_trans_quats_to_pointmap(rays, depth, t1[:b], q[:b])
    _assert_close(G.pointmap_from_rays_depth_pose(rays, depth, t1[:b], q[:b]), pts, 1e-5, "pointmap decode")
    out["depth"], out["pts_world"] = depth.numpy(), pts.numpy()

    dz = depth.clone()
    dz[0, :10] = 0
    nd, nf = RG.normalize_depth_using_non_zero_pixels(dz, return_norm_factor=True)
    md, mf = G.normalize_depth_nonzero(dz)
    _assert_close(md, nd, 1e-6, "normalize_depth_using_non_zero_pixels")
    _assert_close(mf, nf, 1e-6, "  ... factor")
    out["depth_sparse"], out["depth_norm"], out["depth_factor"] = dz.numpy(), nd.numpy(), nf.numpy()
    tv = torch.randn(2, 5, 3, generator=g)
    tv[:, 0] = 0
    nt, ntf = RG.normalize_pose_translations(tv, return_norm_factor=True)
    mt_, mtf = G.normalize_pose_translations(tv)
    _assert_close(mt_, nt, 1e-6, "normalize_pose_translations")
    _assert_close(mtf, ntf, 1e-6, "  ... factor")
    out["trans_views"], out["trans_norm"], out["trans_factor"] = tv.numpy(), nt.numpy(), ntf.numpy()
    lg = RG.apply_log_to_norm(dz)
    _assert_close(G.log_of_norm(dz), lg, 1e-6, "apply_log_to_norm")
    out["depth_log"] = lg.numpy()

    # edge masks on a synthetic scene: tilted plane + a raised box (depth step) + noise + an invalid region
    hh, ww = 60, 72
    ys, xs = np.meshgrid(np.arange(hh, dtype=np.float32), np.arange(ww, dtype=np.float32), indexing="ij")
    z = 2.0 + 0.01 * xs + 0.004 * ys
    z[20:40, 25:50] -= 0.6
    rng = np.random.default_rng(5)
    z = (z + rng.normal(0, 0.002, z.shape)).astype(np.float32)
    pts_np = np.stack([(xs - 36) / 60 * z, (ys - 30) / 60 * z, z], -1).astype(np.float32)
    mask = np.ones((hh, ww), dtype=bool)
    mask[5:12, 60:70] = False
    mask[rng.random((hh, ww)) < 0.02] = False
    normals, nmask = RG.points_to_normals(pts_np, mask=mask)
    mn, mm = G.points_to_normals(pts_np, mask)
    _assert_close(mn, normals, 1e-6, "points_to_normals")
    assert (mm == nmask).all()
    ne = RG.normals_edge(normals, tol=5.0, mask=nmask)
    assert (G.normals_edge(normals, 5.0, nmask) == ne).all(), "normals_edge mismatch"
    de = RG.depth_edge(z, rtol=0.03, mask=mask)
    assert (G.depth_edge(z, 0.03, mask) == de).all(), "depth_edge mismatch"
    print(f"  oracle vs reference  edge masks identical (normal edges {ne.sum()}, depth edges {de.sum()})")
    out.update(edge_pts=pts_np, edge_mask=mask, edge_normals=normals, edge_nmask=nmask, edge_ne=ne, edge_de=de)
    np.savez_compressed(OUT / "geometry.npz", **out)


def golden_inference():
    import mapanything.utils.inference as RI

    import oracle.inference as I

    g = torch.Generator().manual_seed(11)
    out = {}
    b, h, w = 1, 56, 70
    # --- preprocess: intrinsics + depth_z + 4x4 poses; and ray_directions + tuple poses
    k = torch.tensor([[[61.0, 0, 34.2], [0, 59.5, 27.1], [0, 0, 1]]])
    q = torch.randn(1, 4, generator=g)
    q = q / q.norm()
    import mapanything.utils.geometry as RG

    pose = torch.eye(4)[None].clone()
    pose[:, :3, :3] = RG.quaternion_to_rotation_matrix(q)
    pose[:, :3, 3] = torch.tensor([0.2, -0.1, 0.4])
    img = torch.randn(b, 3, h, w, generator=g)
    depth_z = torch.rand(b, h, w, 1, generator=g) * 2 + 1
    rays_in = torch.randn(b, h, w, 3, generator=g) * 0.3 + torch.tensor([0.0, 0.0, 1.0])

    def views():
        return [
            {"img": img.clone(), "data_norm_type": ["dinov2"], "intrinsics": k.clone(), "depth_z": depth_z.clone(),
             "camera_poses": pose.clone()},
            {"img": img.clone(), "data_norm_type": ["dinov2"], "ray_directions": rays_in.clone(),
             "camera_poses": (q.clone(), torch.tensor([[0.5, 0.5, 0.5]])), "is_metric_scale": torch.tensor([False])},
        ]

    rp = RI.preprocess_input_views_for_inference(RI.validate_input_views_for_inference(views()))
    mp = I.preprocess_views(I.validate_views(views()))
    for i in range(2):
        assert set(rp[i].keys()) == set(mp[i].keys()), (rp[i].keys(), mp[i].keys())
        for key in rp[i]:
            if torch.is_tensor(rp[i][key]):
                _assert_close(mp[i][key].float(), rp[i][key].float(), 1e-5, f"preprocess view{i} {key}")
                out[f"pre{i}_{key}"] = rp[i][key].numpy()
    out.update(pre_img=img.numpy(), pre_K=k.numpy(), pre_depth_z=depth_z.numpy(), pre_pose=pose.numpy(), pre_q=q.numpy(),
               pre_rays_in=rays_in.numpy())

    # --- postprocess on synthetic raw outputs (a decodable scene so the edge masks are non-trivial)
    hh, ww = 60, 72
    ys, xs = np.meshgrid(np.arange(hh, dtype=np.float32), np.arange(ww, dtype=np.float32), indexing="ij")
    z = 2.0 + 0.01 * xs + 0.004 * ys
    z[20:40, 25:50] -= 0.6
    rng = np.random.default_rng(9)
    z = (z + rng.normal(0, 0.002, z.shape)).astype(np.float32)
    kk = torch.tensor([[[60.0, 0, 36.0], [0, 60.0, 30.0], [0, 0, 1]]])
    _, rays = RG.get_rays_in_camera_frame(kk, hh, ww, True)
    depth_z2 = torch.from_numpy(z)[None, ..., None]
    dar = torch.norm(depth_z2 * rays / rays[..., 2:3], dim=-1, keepdim=True)
    quat = torch.tensor([[0.05, -0.02, 0.01, 0.99]])
    quat = quat / quat.norm()
    trans = torch.tensor([[0.3, 0.0, -0.1]])
    scale = torch.tensor([[1.7]])
    pts_world = RG.convert_ray_dirs_depth_along_ray_pose_trans_quats_to_pointmap(rays, dar, trans, quat)
    logits = torch.from_numpy(rng.normal(1.0, 1.0, (1, hh, ww)).astype(np.float32))
    conf = 1 + torch.exp(torch.from_numpy(rng.normal(0, 1, (1, hh, ww)).astype(np.float32)))
    raw = [{
        "pts3d": pts_world * scale[..., None, None], "pts3d_cam": rays * dar * scale[..., None, None], "ray_directions": rays,
        "depth_along_ray": dar * scale[..., None, None], "cam_trans": trans * scale, "cam_quats": quat,
        "metric_scaling_factor": scale, "conf": conf, "non_ambiguous_mask": torch.sigmoid(logits) > 0.5,
        "non_ambiguous_mask_logits": logits,
    }]
    img2 = torch.randn(1, 3, hh, ww, generator=g)
    inp = [{"img": img2, "data_norm_type": ["dinov2"]}]
    for tag, kw in [("default", {}), ("conf", {"apply_confidence_mask": True, "confidence_percentile": 25}),
                    ("noedge", {"mask_edges": False})]:
        rpost = RI.postprocess_model_outputs_for_inference([dict(raw[0])], inp, **kw)[0]
        mpost = I.postprocess_outputs([dict(raw[0])], inp, **kw)[0]
        assert set(rpost.keys()) == set(mpost.keys())
        for key in rpost:
            if rpost[key].dtype == torch.bool:
                assert (rpost[key] == mpost[key]).all(), f"postprocess[{tag}] {key}"
            else:
                _assert_close(mpost[key], rpost[key], 1e-3 if key == "intrinsics" else 1e-6, f"postprocess[{tag}] {key}")
            out[f"post_{tag}_{key}"] = rpost[key].numpy()
    for key, val in raw[0].items():
        out[f"raw_{key}"] = val.numpy()
    out["raw_img"] = img2.numpy()
    np.savez_compressed(OUT / "inference.npz", **out)


IMAGE_CASES = {
    # name: (source sizes (W, H), load_images kwargs)
    "fixed": ([(200, 150), (180, 135), (240, 180)], dict(resize_mode="fixed_size", size=(140, 98))),
    "square": ([(90, 160), (100, 150)], dict(resize_mode="square", size=112)),
    "longest": ([(64, 48), (80, 52)], dict(resize_mode="longest_side", size=154)),
    "mapping": ([(320, 240), (300, 235)], dict(resize_mode="fixed_mapping")),
}


def synth_image(w: int, h: int, seed: int) -> np.ndarray:
    """Smooth colour gradients + blobs + noise, (h, w, 3) uint8 -- stands in for a photograph."""
    rng = np.random.default_rng(seed)
    ys, xs = np.meshgrid(np.linspace(0, 1, h), np.linspace(0, 1, w), indexing="ij")
    img = np.stack([xs, ys, 0.5 + 0.5 * np.sin(6 * xs + 4 * ys)], -1) * 200
    for _ in range(4):
        cx, cy, r = rng.uniform(0, 1), rng.uniform(0, 1), rng.uniform(0.05, 0.3)
        img += (((xs - cx) ** 2 + (ys - cy) ** 2) < r * r)[..., None] * rng.uniform(-80, 80, 3)
    img += rng.normal(0, 12, img.shape)
    return np.clip(img, 0, 255).astype(np.uint8)


def golden_image():
    """load_images (mapanything/utils/image.py:134-332) on synthetic PNG files written to a temporary folder."""
    import tempfile

    import PIL.Image
    from mapanything.utils import image as RIM

    from oracle import image as OI

    out = {}
    for ci, (name, (sizes, kw)) in enumerate(IMAGE_CASES.items()):
        with tempfile.TemporaryDirectory() as d:
            for i, (w, h) in enumerate(sizes):
                src = synth_image(w, h, 100 * ci + i)
                PIL.Image.fromarray(src).save(f"{d}/img_{i:02d}.png")
                out[f"{name}_src{i}"] = src
            (Path(d) / "notes.txt").write_text("not an image")  # skipped by extension
            ref = RIM.load_images(d, **kw)
            mine = OI.load_images(d, **kw)
        assert len(ref) == len(mine) == len(sizes)
        for i, (rv, mv) in enumerate(zip(ref, mine)):
            assert set(rv.keys()) == set(mv.keys())
            assert rv["img"].dtype == torch.float32 and rv["img"].shape == mv["img"].shape, (rv["img"].shape, mv["img"].shape)
            assert torch.equal(rv["img"], mv["img"]), f"load_images[{name}] view {i}: oracle differs from the reference"
            assert (rv["true_shape"] == mv["true_shape"]).all() and rv["idx"] == mv["idx"] and rv["instance"] == mv["instance"]
            assert rv["data_norm_type"] == mv["data_norm_type"]
            img = rv["img"].numpy()
            if name != "fixed":  # keep the fixture small: a strided sample and exact per-channel sums
                st = 7 if name == "mapping" else 3
                out[f"{name}_img{i}_sample"] = img[:, :, ::st, ::st].copy()
                out[f"{name}_img{i}_stride"] = np.array(st)
                out[f"{name}_img{i}_sum"] = img.astype(np.float64).sum(axis=(0, 2, 3))
                out[f"{name}_img{i}_shape"] = np.array(img.shape)
            else:
                out[f"{name}_img{i}"] = img
            out[f"{name}_true_shape{i}"] = rv["true_shape"]
        print(f"  oracle vs reference  load_images[{name:<8s}] {tuple(ref[0]['img'].shape)}            bit-exact")
    np.savez_compressed(OUT / "image.npz", **out)


PREPROCESS_CASES = {
    "fixed": dict(resize_mode="fixed_size", size=(140, 98)),
    "square": dict(resize_mode="square", size=112),
    "longest": dict(resize_mode="longest_side", size=154),
}


def synth_multimodal_views(seed: int = 5):
    """Four views covering every input form preprocess_inputs accepts (image.py:345-356): uint8 / float arrays, float
    tensors, PIL images; intrinsics (float32 array, tensor, float64), ray directions, depth (array / tensor), poses as a
    matrix and as a (quats, trans) tuple, pass-through keys."""
    import PIL.Image

    from oracle.geometry import rays_from_intrinsics

    rng = np.random.default_rng(seed)
    views = []
    W, H = 200, 150
    views.append(dict(img=synth_image(W, H, 1), intrinsics=np.array([[180.0, 0, 101.3], [0, 182.0, 73.9], [0, 0, 1]], np.float32),
                      depth_z=rng.uniform(1, 4, (H, W)).astype(np.float32), camera_poses=np.eye(4, dtype=np.float32),
                      is_metric_scale=torch.tensor([True])))
    W, H = 180, 135
    views.append(dict(img=torch.from_numpy(synth_image(W, H, 2)).float() / 255,
                      intrinsics=torch.tensor([[170.0, 0, 88.0], [0, 170.0, 70.0], [0, 0, 1]]),
                      camera_poses=(torch.tensor([0, 0, 0, 1.0]), np.array([0.1, 0.2, 0.3], np.float32))))
    W, H = 240, 181
    rays = rays_from_intrinsics(torch.tensor([[[210.0, 0, 119.5], [0, 209.0, 90.2], [0, 0, 1]]]), H, W)[0]
    views.append(dict(img=PIL.Image.fromarray(synth_image(W, H, 3)), ray_directions=rays,
                      depth_z=torch.from_numpy(rng.uniform(1, 4, (H, W)).astype(np.float32))))
    views.append(dict(img=synth_image(161, 120, 4).astype(np.float32) / 255.0, instance="x"))
    W, H = 210, 140
    views.append(dict(img=torch.from_numpy(synth_image(W, H, 6)).float(),  # float tensor in [0, 255]
                      intrinsics=np.array([[190.0, 0, 104.5], [0, 190.0, 69.5], [0, 0, 1]], np.float64),
                      depth_z=rng.uniform(0, 3, (H, W)).astype(np.float32)))
    return views


def golden_preprocess_inputs():
    """preprocess_inputs (mapanything/utils/image.py:335-675) on synthetic multi-modal views."""
    from mapanything.utils import geometry as RG
    from mapanything.utils import image as RIM

    from oracle import image as OI

    out = {}
    for name, kw in PREPROCESS_CASES.items():
        ref = RIM.preprocess_inputs(synth_multimodal_views(), **kw)
        mine = OI.preprocess_inputs(synth_multimodal_views(), **kw)
        for i, (rv, mv) in enumerate(zip(ref, mine)):
            assert set(rv.keys()) == set(mv.keys()), (rv.keys(), mv.keys())
            for key in rv:
                if key == "camera_poses" and isinstance(rv[key], tuple):
                    assert all(torch.equal(a, b) for a, b in zip(rv[key], mv[key]))
                    out[f"{name}_v{i}_pose_q"], out[f"{name}_v{i}_pose_t"] = rv[key][0].numpy(), rv[key][1].numpy()
                elif torch.is_tensor(rv[key]):
                    assert rv[key].dtype == mv[key].dtype and rv[key].shape == mv[key].shape, (name, i, key)
                    if "ray_directions" in synth_multimodal_views()[i] and key == "intrinsics":
                        # recovered from rays by least squares: the oracle's own solver, tolerance instead of bits
                        _assert_close(mv[key], rv[key], 2e-3, f"preprocess_inputs[{name}] v{i} intrinsics (from rays)")
                    else:
                        assert torch.equal(rv[key], mv[key]), f"preprocess_inputs[{name}] view {i} {key}: oracle differs"
                    arr = rv[key].numpy()
                    if key == "img" and name != "fixed":  # keep the fixture small: strided sample + exact sums
                        out[f"{name}_v{i}_img_sample"] = arr[:, :, ::3, ::3].copy()
                        out[f"{name}_v{i}_img_sum"] = arr.astype(np.float64).sum(axis=(0, 2, 3))
                        out[f"{name}_v{i}_img_shape"] = np.array(arr.shape)
                    else:
                        out[f"{name}_v{i}_{key}"] = arr
                else:
                    assert rv[key] == mv[key]
        print(f"  oracle vs reference  preprocess_inputs[{name:<8s}] {len(ref)} views                      bit-exact")
    np.savez_compressed(OUT / "preprocess_inputs.npz", **out)


def golden_depthmap():
    """depthmap_to_camera_frame / depthmap_to_world_frame (mapanything/utils/geometry.py:18-114)."""
    from mapanything.utils import geometry as RG

    from oracle import geometry as G

    g = torch.Generator().manual_seed(21)
    b, h, w = 2, 24, 32
    depth = torch.rand(b, h, w, generator=g) * 3
    depth[0, 3:6, 4:9] = 0.0  # invalid pixels
    K = torch.tensor([[[30.0, 0, 15.5], [0, 31.0, 11.5], [0, 0, 1]], [[28.0, 0, 16.2], [0, 28.5, 12.3], [0, 0, 1]]])
    q = torch.randn(b, 4, generator=g)
    q = q / q.norm(dim=-1, keepdim=True)
    pose = torch.eye(4).repeat(b, 1, 1)
    pose[:, :3, :3] = RG.quaternion_to_rotation_matrix(q)
    pose[:, :3, 3] = torch.randn(b, 3, generator=g)
    out = dict(depth=depth.numpy(), K=K.numpy(), pose=pose.numpy())
    rc, rv = RG.depthmap_to_camera_frame(depth, K)
    mc, mv = G.depthmap_to_camera_frame(depth, K)
    assert torch.equal(rc, mc) and torch.equal(rv, mv)
    rw, rv2 = RG.depthmap_to_world_frame(depth, K, pose)
    mw, _ = G.depthmap_to_world_frame(depth, K, pose)
    _assert_close(mw, rw, 1e-6, "depthmap_to_world_frame")
    r1, v1 = RG.depthmap_to_world_frame(depth[0], K[0], pose[0])  # un-batched form
    assert torch.equal(r1, rw[0]) or (r1 - rw[0]).abs().max() < 1e-6
    out.update(pts_cam=rc.numpy(), valid=rv.numpy(), pts_world=rw.numpy())
    np.savez_compressed(OUT / "depthmap.npz", **out)


def main():
    assert REF.exists(), "this script needs /root/reference (build container only)"
    _install_reference_stubs()
    OUT.mkdir(parents=True, exist_ok=True)
    torch.manual_seed(0)
    print("geometry:")
    golden_geometry()
    print("inference pre/post:")
    golden_inference()
    print("depthmap:")
    golden_depthmap()
    print("load_images:")
    golden_image()
    print("preprocess_inputs:")
    golden_preprocess_inputs()
    print("DINOv2 ViT:")
    golden_vit()
    for f in sorted(OUT.glob("*.npz")):
        print(f"wrote {f} ({f.stat().st_size / 1024:.0f} KiB)")


if __name__ == "__main__":
    main()
