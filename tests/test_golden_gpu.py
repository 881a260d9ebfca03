"""GPU parity tests of the CUDA post-processing / decode kernels against the golden vectors generated from the REFERENCE's own
code (oracle/make_golden.py -> tests/golden/*.npz): the same fixtures tests/test_oracle_golden.py pins the oracle with, fed
to the product path (`mapanything_b200.inference`, `ops.decode_dense`) on the device.

Bar: bit-exact for masks and every boolean, 1e-6 for floats (1e-3 for the least-squares intrinsics), as in the CPU tests.
"""
from pathlib import Path

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

GOLD = Path(__file__).parent / "golden"


def _c(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


@pytest.mark.parametrize("tag,kw", [("default", {}), ("conf", {"apply_confidence_mask": True, "confidence_percentile": 25}),
                                    ("noedge", {"mask_edges": False})])
def test_postprocess_matches_reference_golden(tag, kw):
    """reference mapanything/utils/inference.py:294-480 on its own raw outputs == the CUDA post-processing, incl. the edge
    masks (utils/geometry.py normals_edge / depth_edge / points_to_normals) bit for bit."""
    from mapanything_b200.inference import postprocess_model_outputs_for_inference

    g = np.load(GOLD / "inference.npz")
    raw = {k[4:]: _c(g[k]) for k in g.files if k.startswith("raw_") and k != "raw_img"}
    view = {"img": _c(g["raw_img"]), "data_norm_type": ["dinov2"]}
    out = postprocess_model_outputs_for_inference([raw], [view], **kw)[0]
    torch.cuda.synchronize()
    keys = [k[len(f"post_{tag}_"):] for k in g.files if k.startswith(f"post_{tag}_")]
    assert set(keys) == set(out.keys())
    for key in keys:
        ref = torch.from_numpy(g[f"post_{tag}_{key}"])
        got = out[key].cpu()
        assert got.shape == ref.shape and got.dtype == ref.dtype, (key, got.shape, ref.shape, got.dtype, ref.dtype)
        if ref.dtype == torch.bool:
            assert torch.equal(got, ref), f"{tag}/{key}: {(got != ref).sum().item()} of {ref.numel()} differ"
        else:
            assert torch.allclose(got, ref, atol=1e-3 if key == "intrinsics" else 1e-6), \
                f"{tag}/{key}: max abs diff {(got - ref).abs().max().item()}"
    m = out["mask"]
    assert 0 < m.sum().item() < m.numel()


def test_edge_mask_kernel_matches_reference_golden():
    """ma_edge_mask on the reference's edge fixture: mask & ~(depth_edge & normals_edge), bit-exact (the reference pairs each
    neighbour angle with the transposed mask window; the kernel reproduces that)."""
    from mapanything_b200.inference import edge_mask

    geo = np.load(GOLD / "geometry.npz")
    pts = _c(geo["edge_pts"][None])
    m = _c(geo["edge_mask"][None])
    got = edge_mask(pts, pts, m, 5.0, 0.03).cpu().numpy()[0]
    want = geo["edge_mask"] & ~(geo["edge_de"] & geo["edge_ne"])
    assert np.array_equal(got, want), f"{(got != want).sum()} of {want.size} pixels differ"
    assert (geo["edge_de"] & geo["edge_ne"] & geo["edge_mask"]).sum() > 0  # the fixture removes something


def test_intrinsics_from_rays_matches_reference_golden():
    """ma_intrinsics_from_rays vs recover_pinhole_intrinsics_from_ray_directions (utils/geometry.py) outputs."""
    from mapanything_b200.inference import intrinsics_from_rays

    geo = np.load(GOLD / "geometry.npz")
    K = intrinsics_from_rays(_c(geo["rays"])).cpu()
    ref = torch.from_numpy(geo["K_rec"])
    assert torch.allclose(K, ref, atol=1e-3), (K - ref).abs().max()


def test_decode_dense_matches_reference_pointmap_golden():
    """ma_decode_dense (dense adaptor + pose adaptor + convert_ray_dirs_depth_along_ray_pose_trans_quats_to_pointmap,
    utils/geometry.py:855-907) on raw values chosen so that the adaptors reproduce the golden inputs: unit rays stay unit
    rays, raw depth = log(depth), unit quaternions, scale logit 0."""
    from mapanything_b200 import ops

    geo = np.load(GOLD / "geometry.npz")
    rays, depth = torch.from_numpy(geo["rays"]), torch.from_numpy(geo["depth"])
    n, H, W, _ = rays.shape
    q = torch.from_numpy(geo["q"])[:n]
    t = torch.from_numpy(geo["t1"])[:n]
    raw = torch.zeros(n * H * W, 8)
    raw[:, :3] = rays.reshape(-1, 3)
    raw[:, 3] = depth.reshape(-1).log()
    raw[:, 4] = 0.25   # conf = 1 + exp(0.25)
    raw[:, 5] = torch.linspace(-2, 2, n * H * W)  # mask logits
    pose_raw = torch.cat([t, q], dim=1).contiguous()
    out = ops.decode_dense(raw.cuda(), pose_raw.cuda(), torch.zeros(1).cuda(), n, H, W)
    torch.cuda.synchronize()
    pts_ref = torch.from_numpy(geo["pts_world"])
    got = out["pts3d"].cpu()
    # exp(log(d)) costs ~2 ulp of the depth; everything else is the reference's own arithmetic
    err = ((got - pts_ref).abs() / pts_ref.abs().clamp(min=1.0)).max().item()
    assert err < 2e-6, err
    assert torch.allclose(out["ray_directions"].cpu(), rays, atol=1e-6)
    assert torch.allclose(out["depth_along_ray"].cpu(), depth, rtol=1e-6, atol=0)
    qn = q / q.norm(dim=1, keepdim=True)
    assert torch.allclose(out["cam_quats"].cpu(), qn, atol=1e-6)
    assert torch.equal(out["cam_trans"].cpu(), t)
    assert out["metric_scaling_factor"].item() == 1.0
    assert torch.allclose(out["conf"].cpu(), torch.full((n, H, W), 1.0 + float(np.exp(np.float32(0.25)))), atol=1e-6)
    logits = raw[:, 5].reshape(n, H, W)
    assert torch.equal(out["non_ambiguous_mask_logits"].cpu(), logits)
    assert torch.equal(out["non_ambiguous_mask"].cpu(), torch.sigmoid(logits) > 0.5)
    assert torch.allclose(out["pts3d_cam"].cpu(), rays * depth, rtol=2e-6, atol=1e-7)
