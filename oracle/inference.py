"""Oracle restatement of the infer() pre/post-processing (test infrastructure only).

Follows /root/reference/mapanything/utils/inference.py:
  validate_views      <- validate_input_views_for_inference   :128-199
  preprocess_views    <- preprocess_input_views_for_inference :202-291
  postprocess_outputs <- postprocess_model_outputs_for_inference :294-480 (+ rgb, mapanything/utils/image.py:93-131)
"""
from __future__ import annotations

from typing import Any, Dict, List

import numpy as np
import torch

from . import geometry as G

ALLOWED_VIEW_KEYS = {
    "img", "data_norm_type", "intrinsics", "ray_directions", "depth_z", "camera_poses", "is_metric_scale",
    "instance", "idx", "true_shape",
}
REQUIRED_KEYS = {"img", "data_norm_type"}
CONFLICTING_KEYS = [("intrinsics", "ray_directions")]

# uniception IMAGE_NORMALIZATION_DICT["dinov2"] (ImageNet statistics; cf. vggt/models/aggregator.py:23-24)
IMAGE_NORMALIZATION = {
    "dinov2": ((0.485, 0.456, 0.406), (0.229, 0.224, 0.225)),
    "identity": ((0.0, 0.0, 0.0), (1.0, 1.0, 1.0)),
    "dust3r": ((0.5, 0.5, 0.5), (0.5, 0.5, 0.5)),
}


def validate_views(views: List[Dict[str, Any]]) -> List[Dict[str, Any]]:
    if not views:
        raise ValueError("At least one view must be provided")
    with_pose = []
    for i, view in enumerate(views):
        keys = set(view.keys())
        bad = keys - ALLOWED_VIEW_KEYS
        if bad:
            raise ValueError(f"View {i} contains invalid keys: {bad}. Allowed keys are: {sorted(ALLOWED_VIEW_KEYS)}")
        missing = REQUIRED_KEYS - keys
        if missing:
            raise ValueError(f"View {i} missing required keys: {missing}")
        for group in CONFLICTING_KEYS:
            present = [k for k in group if k in keys]
            if len(present) > 1:
                raise ValueError(
                    f"View {i} contains conflicting keys: {present}. Only one of {group} can be provided at a time."
                )
        if "depth_z" in keys and "intrinsics" not in keys and "ray_directions" not in keys:
            raise ValueError(
                f"View {i} depth constraint violation: If 'depth_z' is provided, then 'intrinsics' or "
                f"'ray_directions' must also be provided."
            )
        if "camera_poses" in keys:
            with_pose.append(i)
    if with_pose and 0 not in with_pose:
        raise ValueError(
            f"Camera pose constraint violation: Views {with_pose} have camera_poses, but view 0 (reference view) does not."
        )
    return views


def preprocess_views(views: List[Dict[str, Any]]) -> List[Dict[str, Any]]:
    out = []
    for i, view in enumerate(views):
        pv = dict(view)
        if "intrinsics" in view:
            h, w = view["img"].shape[-2:]
            pv["ray_directions"] = G.rays_from_intrinsics(view["intrinsics"], h, w, unit_sphere=True)
            del pv["intrinsics"]
        elif "ray_directions" in view:
            r = view["ray_directions"]
            pv["ray_directions"] = r / (torch.norm(r, dim=-1, keepdim=True) + 1e-8)
        if "depth_z" in view:
            r = pv["ray_directions"]
            pts = view["depth_z"] * (r / r[..., 2:3])
            pv["depth_along_ray"] = torch.norm(pts, dim=-1, keepdim=True)
            del pv["depth_z"]
        if "camera_poses" in view:
            poses = view["camera_poses"]
            if isinstance(poses, tuple) and len(poses) == 2:
                pv["camera_pose_quats"], pv["camera_pose_trans"] = poses
            elif torch.is_tensor(poses) and poses.shape[-2:] == (4, 4):
                pv["camera_pose_quats"] = G.rotmat_to_quat(poses[:, :3, :3])
                pv["camera_pose_trans"] = poses[:, :3, 3]
            else:
                raise ValueError(
                    f"View {i}: camera_poses must be either a tuple of (quats, trans) or a tensor of (B, 4, 4) "
                    f"transformation matrices."
                )
            del pv["camera_poses"]
        if "is_metric_scale" not in pv:
            pv["is_metric_scale"] = torch.ones(view["img"].shape[0], dtype=torch.bool, device=view["img"].device)
        if "ray_directions" in pv:
            pv["ray_directions_cam"] = pv.pop("ray_directions")
        out.append(pv)
    return out


def denormalize_image(img: torch.Tensor, norm_type: str) -> torch.Tensor:
    """image.py:93-131 rgb(): (B,3,H,W) normalised -> (B,H,W,3) in [0,1] (computed through numpy like the reference)."""
    mean, std = IMAGE_NORMALIZATION[norm_type]
    x = img.detach().cpu().permute(0, 2, 3, 1).numpy()
    x = x * np.array(std, dtype=np.float32).reshape(1, 1, 1, 3) + np.array(mean, dtype=np.float32).reshape(1, 1, 1, 3)
    return torch.from_numpy(np.clip(x, 0.0, 1.0)).to(img.device)


def postprocess_outputs(raw_outputs, input_views, apply_mask=True, mask_edges=True, edge_normal_threshold=5.0,
                        edge_depth_threshold=0.03, apply_confidence_mask=False, confidence_percentile=10):
    results = []
    for raw, view in zip(raw_outputs, input_views):
        out = dict(raw)
        img = view["img"]
        out["img_no_norm"] = denormalize_image(img, view["data_norm_type"][0])
        if "pts3d_cam" in out:
            out["depth_z"] = out["pts3d_cam"][..., 2:3]
        if "ray_directions" in out:
            out["intrinsics"] = G.intrinsics_from_rays(out["ray_directions"])
        if "cam_trans" in out and "cam_quats" in out:
            b = out["cam_trans"].shape[0]
            pose = torch.eye(4, device=img.device).unsqueeze(0).repeat(b, 1, 1)
            pose[:, :3, :3] = G.quat_to_rotmat(out["cam_quats"])
            pose[:, :3, 3] = out["cam_trans"]
            out["camera_poses"] = pose
        if apply_mask:
            final = None
            if "non_ambiguous_mask" in out:
                final = out["non_ambiguous_mask"].cpu().numpy()
            if apply_confidence_mask and "conf" in out:
                conf = out["conf"].cpu()
                b = conf.shape[0]
                thr = torch.quantile(conf.reshape(b, -1), confidence_percentile / 100.0, dim=1).view(b, 1, 1)
                cm = (conf > thr).numpy()
                final = cm if final is None else final & cm
            if mask_edges and final is not None and "pts3d" in out:
                pts = out["pts3d"].cpu().numpy()
                edge_masks = []
                for b in range(final.shape[0]):
                    if final[b].any():
                        normals, nmask = G.points_to_normals(pts[b], final[b])
                        ne = G.normals_edge(normals, edge_normal_threshold, nmask)
                        dz = out["depth_z"][b].squeeze(-1).cpu().numpy()
                        de = G.depth_edge(dz, edge_depth_threshold, final[b])
                        edge_masks.append(~(de & ne))
                    else:
                        edge_masks.append(np.zeros_like(final[b], dtype=bool))
                final = final & np.stack(edge_masks, axis=0)
            if final is not None:
                m = torch.from_numpy(final).to(out["pts3d"].device).unsqueeze(-1)
                for key in ("pts3d", "pts3d_cam", "depth_along_ray", "depth_z"):
                    if key in out:
                        out[key] = out[key] * m
                out["mask"] = m
        results.append(out)
    return results
