"""Per-view conversions of preprocess_input_views_for_inference (reference mapanything/utils/inference.py:224-272) on the
GPU: intrinsics -> unit ray directions, ray normalisation, depth_z -> depth along ray, 4x4 poses -> (quats, trans)."""
from __future__ import annotations

from typing import Any, Dict

import torch

from . import _lib
from .ops import _stream, check, launch


def _f32(t: torch.Tensor) -> torch.Tensor:
    if not t.is_cuda:
        raise ValueError("mapanything_b200 preprocessing runs on the GPU: move the views to the model's device (infer() does)")
    return t.to(torch.float32).contiguous()


def rays_from_intrinsics(intrinsics: torch.Tensor, height: int, width: int) -> torch.Tensor:
    k = _f32(intrinsics)
    if k.dim() == 2:
        k = k[None]
    out = torch.empty(k.shape[0], height, width, 3, device=k.device, dtype=torch.float32)
    with launch("preprocess"):
        check(_lib.load().ma_rays_from_intrinsics(k.data_ptr(), out.data_ptr(), k.shape[0], height, width, _stream()),
              "ma_rays_from_intrinsics")
    return out


def normalize_rays(rays: torch.Tensor) -> torch.Tensor:
    r = _f32(rays)
    out = torch.empty_like(r)
    with launch("preprocess"):
        check(_lib.load().ma_normalize_rays(r.data_ptr(), out.data_ptr(), r.numel() // 3, _stream()), "ma_normalize_rays")
    return out


def depth_z_to_along_ray(depth_z: torch.Tensor, rays: torch.Tensor) -> torch.Tensor:
    d, r = _f32(depth_z), _f32(rays)
    if d.numel() * 3 != r.numel():
        raise ValueError(f"depth_z {tuple(depth_z.shape)} does not match ray_directions {tuple(rays.shape)}")
    out = torch.empty(*r.shape[:-1], 1, device=d.device, dtype=torch.float32)
    with launch("preprocess"):
        check(_lib.load().ma_depth_z_to_along_ray(d.data_ptr(), r.data_ptr(), out.data_ptr(), d.numel(), _stream()),
              "ma_depth_z_to_along_ray")
    return out


def pose_to_quat_trans(poses: torch.Tensor):
    p = _f32(poses)
    b = p.shape[0]
    q = torch.empty(b, 4, device=p.device, dtype=torch.float32)
    t = torch.empty(b, 3, device=p.device, dtype=torch.float32)
    with launch("preprocess"):
        check(_lib.load().ma_pose_to_quat_trans(p.data_ptr(), q.data_ptr(), t.data_ptr(), b, _stream()), "ma_pose_to_quat_trans")
    return q, t


def convert_view(pv: Dict[str, Any], view: Dict[str, Any], view_idx: int) -> Dict[str, Any]:
    """Steps 1-3 of the reference function for one view (pv is the shallow copy being built)."""
    if "intrinsics" in view:
        height, width = view["img"].shape[-2:]
        pv["ray_directions"] = rays_from_intrinsics(view["intrinsics"], height, width)
        del pv["intrinsics"]
    elif "ray_directions" in view:
        pv["ray_directions"] = normalize_rays(view["ray_directions"])
    if "depth_z" in view:
        pv["depth_along_ray"] = depth_z_to_along_ray(view["depth_z"], pv["ray_directions"])
        del pv["depth_z"]
    if "camera_poses" in view:
        cp = view["camera_poses"]
        if isinstance(cp, tuple) and len(cp) == 2:
            pv["camera_pose_quats"], pv["camera_pose_trans"] = cp
        elif torch.is_tensor(cp) and cp.shape[-2:] == (4, 4):
            pv["camera_pose_quats"], pv["camera_pose_trans"] = pose_to_quat_trans(cp)
        else:
            raise ValueError(
                f"View {view_idx}: camera_poses must be either a tuple of (quats, trans) "
                f"or a tensor of (B, 4, 4) transformation matrices."
            )
        del pv["camera_poses"]
    return pv
