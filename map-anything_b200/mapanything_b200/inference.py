"""Host-side mirror of the reference's infer() helpers (mapanything/utils/inference.py), same names, argument
meaning and error behaviour, with the arithmetic moved to the GPU kernels:

  validate_input_views_for_inference        :128-199  pure Python checks (identical messages / ValueError)
  preprocess_input_views_for_inference      :202-291  key conversion; per-pixel math in ma_* kernels
  postprocess_model_outputs_for_inference   :294-480  denorm / intrinsics / poses / edge masks on the device
"""
from __future__ import annotations

import ctypes as C
from typing import Any, Dict, List

import torch

from . import _lib, ops
from ._lib import check

ALLOWED_VIEW_KEYS = {
    "img", "data_norm_type", "depth_z", "ray_directions", "intrinsics", "camera_poses", "is_metric_scale", "true_shape",
    "idx", "instance",
}
REQUIRED_KEYS = {"img", "data_norm_type"}
CONFLICTING_KEYS = [("intrinsics", "ray_directions")]

# uniception.models.encoders.image_normalizations.IMAGE_NORMALIZATION_DICT (values used by image.py:93-131 rgb())
IMAGE_NORMALIZATION_DICT = {
    "dinov2": ((0.485, 0.456, 0.406), (0.229, 0.224, 0.225)),
    "dust3r": ((0.5, 0.5, 0.5), (0.5, 0.5, 0.5)),
    "identity": ((0.0, 0.0, 0.0), (1.0, 1.0, 1.0)),
}


def validate_input_views_for_inference(views: List[Dict[str, Any]], first_view_is_reference: bool = True) -> List[Dict[str, Any]]:
    """reference utils/inference.py:128-199.  first_view_is_reference=False (ranks > 0 of a view-sharded scene, whose first
    local view is not the scene's view 0) skips the reference-view pose rule, which is then checked scene-wide."""
    if not views:
        raise ValueError("At least one view must be provided")
    views_with_poses = []
    for view_idx, view in enumerate(views):
        provided_keys = set(view.keys())
        invalid_keys = provided_keys - ALLOWED_VIEW_KEYS
        if invalid_keys:
            raise ValueError(
                f"View {view_idx} contains invalid keys: {invalid_keys}. Allowed keys are: {sorted(ALLOWED_VIEW_KEYS)}"
            )
        missing_keys = REQUIRED_KEYS - provided_keys
        if missing_keys:
            raise ValueError(f"View {view_idx} missing required keys: {missing_keys}")
        for conflict_set in CONFLICTING_KEYS:
            present = [k for k in conflict_set if k in provided_keys]
            if len(present) > 1:
                raise ValueError(
                    f"View {view_idx} contains conflicting keys: {present}. "
                    f"Only one of {conflict_set} can be provided at a time."
                )
        if "depth_z" in provided_keys and "intrinsics" not in provided_keys and "ray_directions" not in provided_keys:
            raise ValueError(
                f"View {view_idx} depth constraint violation: If 'depth_z' is provided, "
                f"then 'intrinsics' or 'ray_directions' must also be provided. "
                f"Z Depth values require camera calibration information to be meaningful for an image."
            )
        if "camera_poses" in provided_keys:
            views_with_poses.append(view_idx)
    if first_view_is_reference and views_with_poses and 0 not in views_with_poses:
        raise ValueError(
            f"Camera pose constraint violation: Views {views_with_poses} have camera_poses, "
            f"but view 0 (reference view) does not. When using camera_poses, the first view "
            f"must also provide camera_poses to serve as the reference frame."
        )
    return views


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def denorm_image(img: torch.Tensor, norm_type: str) -> torch.Tensor:
    if norm_type not in IMAGE_NORMALIZATION_DICT:
        raise ValueError(
            f"Unknown image normalization type: {norm_type}. Available types: identity or {IMAGE_NORMALIZATION_DICT.keys()}"
        )
    mean, std = IMAGE_NORMALIZATION_DICT[norm_type]
    img = img.contiguous().float()
    n, _, H, W = img.shape
    out = torch.empty(n, H, W, 3, device=img.device, dtype=torch.float32)
    m3, s3 = (C.c_float * 3)(*mean), (C.c_float * 3)(*std)
    check(_lib.load().ma_denorm_image(img.data_ptr(), out.data_ptr(), n, H, W, m3, s3, _stream()), "ma_denorm_image")
    ops._count()
    return out


def intrinsics_from_rays(rays: torch.Tensor) -> torch.Tensor:
    rays = rays.contiguous()
    n, H, W, _ = rays.shape
    K = torch.empty(n, 3, 3, device=rays.device, dtype=torch.float32)
    check(_lib.load().ma_intrinsics_from_rays(rays.data_ptr(), K.data_ptr(), n, H, W, _stream()), "ma_intrinsics_from_rays")
    ops._count()
    return K


def pose_matrices(quats: torch.Tensor, trans: torch.Tensor) -> torch.Tensor:
    n = quats.shape[0]
    out = torch.empty(n, 4, 4, device=quats.device, dtype=torch.float32)
    check(_lib.load().ma_pose_matrices(quats.contiguous().data_ptr(), trans.contiguous().data_ptr(), out.data_ptr(), n, _stream()),
          "ma_pose_matrices")
    ops._count()
    return out


def edge_mask(pts3d: torch.Tensor, pts3d_cam: torch.Tensor, mask_in: torch.Tensor, normal_tol_deg: float, depth_rtol: float):
    """mask_in (n,H,W) bool -> mask_in & ~(depth_edge & normal_edge), (n,H,W) bool."""
    n, H, W, _ = pts3d.shape
    dev = pts3d.device
    pts3d, pts3d_cam = pts3d.contiguous(), pts3d_cam.contiguous()
    mask_in = mask_in.contiguous()
    out = torch.empty(n, H, W, device=dev, dtype=torch.bool)
    ws_n = torch.empty(n, H, W, 3, device=dev, dtype=torch.float32)
    ws_m = torch.empty(n, H, W, device=dev, dtype=torch.uint8)
    ws_a = torch.empty(n, H, W, device=dev, dtype=torch.float32)
    ws_d = torch.empty(n, H, W, device=dev, dtype=torch.uint8)
    check(
        _lib.load().ma_edge_mask(
            pts3d.data_ptr(), pts3d_cam.data_ptr() + 8, 3, mask_in.data_ptr(), out.data_ptr(), ws_n.data_ptr(), ws_m.data_ptr(),
            ws_a.data_ptr(), ws_d.data_ptr(), n, H, W, float(normal_tol_deg), float(depth_rtol), _stream(),
        ),
        "ma_edge_mask",
    )
    ops._count(3)
    return out


def mask_dense(x: torch.Tensor, mask: torch.Tensor, width: int, in_stride: int = None, in_offset: int = 0) -> torch.Tensor:
    """x: fp32 (..., C) contiguous; returns a new tensor (pixels, width) = x[..., in_offset:in_offset+width] * mask."""
    x = x.contiguous()
    pixels = mask.numel()
    in_stride = in_stride or x.shape[-1]
    out = torch.empty(*mask.shape, width, device=x.device, dtype=torch.float32)
    check(_lib.load().ma_apply_mask(x.data_ptr(), in_stride, in_offset, mask.data_ptr(), out.data_ptr(), pixels, width, _stream()),
          "ma_apply_mask")
    ops._count()
    return out


def mask_and(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    a, b = a.contiguous(), b.contiguous()
    out = torch.empty_like(a)
    check(_lib.load().ma_mask_and(a.data_ptr(), b.data_ptr(), out.data_ptr(), a.numel(), _stream()), "ma_mask_and")
    ops._count()
    return out


def quantile_mask(conf: torch.Tensor, q: float) -> torch.Tensor:
    """conf (B,H,W) fp32 -> bool (B,H,W): conf > torch.quantile(conf per image, q), computed by ma_quantile_mask."""
    conf = conf.contiguous()
    b = conf.shape[0]
    out = torch.empty(conf.shape, device=conf.device, dtype=torch.bool)
    check(_lib.load().ma_quantile_mask(conf.data_ptr(), out.data_ptr(), None, b, conf.numel() // b, float(q), _stream()),
          "ma_quantile_mask")
    ops._count()
    return out


def postprocess_model_outputs_for_inference(
    raw_outputs: List[Dict[str, torch.Tensor]],
    input_views: List[Dict[str, Any]],
    apply_mask: bool = True,
    mask_edges: bool = True,
    edge_normal_threshold: float = 5.0,
    edge_depth_threshold: float = 0.03,
    apply_confidence_mask: bool = False,
    confidence_percentile: float = 10,
) -> List[Dict[str, torch.Tensor]]:
    """Same outputs as the reference function, computed per view on the device (no D2H, no numpy loop)."""
    dev = next((t.device for r in raw_outputs for t in r.values() if torch.is_tensor(t) and t.is_cuda), None)
    if dev is None:
        raise RuntimeError("postprocess_model_outputs_for_inference: mapanything_b200 has no CPU path (outputs must be CUDA tensors)")
    with torch.cuda.device(dev):
        return _postprocess(raw_outputs, input_views, apply_mask, mask_edges, edge_normal_threshold, edge_depth_threshold,
                            apply_confidence_mask, confidence_percentile)


def _postprocess(raw_outputs, input_views, apply_mask, mask_edges, edge_normal_threshold, edge_depth_threshold,
                 apply_confidence_mask, confidence_percentile):
    processed = []
    for raw, view in zip(raw_outputs, input_views):
        out = dict(raw)
        img = view["img"]
        out["img_no_norm"] = denorm_image(img, view["data_norm_type"][0])
        if "pts3d_cam" in out:
            out["depth_z"] = out["pts3d_cam"][..., 2:3]
        if "ray_directions" in out:
            out["intrinsics"] = intrinsics_from_rays(out["ray_directions"])
        if "cam_trans" in out and "cam_quats" in out:
            out["camera_poses"] = pose_matrices(out["cam_quats"], out["cam_trans"])
        if apply_mask:
            final = out.get("non_ambiguous_mask")
            if apply_confidence_mask and "conf" in out:
                cm = quantile_mask(out["conf"], confidence_percentile / 100.0)
                final = cm if final is None else mask_and(final, cm)
            if mask_edges and final is not None and "pts3d" in out:
                if "pts3d_cam" not in out:   # the reference reads processed_output["depth_z"] here (inference.py:440)
                    raise KeyError("depth_z")
                final = edge_mask(out["pts3d"], out["pts3d_cam"], final, edge_normal_threshold, edge_depth_threshold)
            if final is not None:
                _apply_final_mask(out, final.contiguous())
        processed.append(out)
    return processed


def postprocess_scene(stacked: Dict[str, torch.Tensor], imgs: torch.Tensor, norm_type: str, apply_mask: bool = True,
                      mask_edges: bool = True, edge_normal_threshold: float = 5.0, edge_depth_threshold: float = 0.03,
                      apply_confidence_mask: bool = False, confidence_percentile: float = 10) -> List[Dict[str, torch.Tensor]]:
    """postprocess_model_outputs_for_inference for a whole scene at once: `stacked` holds the forward outputs of all V views
    as [V, ...] tensors (what ma_decode_dense wrote), `imgs` the (V,3,H,W) normalised images.  ONE launch of each kernel
    covers every view -- ~10 launches per scene instead of ~12 per view (100 - 1000-view scenes) -- and the per-view result
    dicts are slices of the scene tensors.  Same arithmetic, kernel for kernel, as the per-view path."""
    with torch.cuda.device(imgs.device):
        V = imgs.shape[0]
        out = dict(stacked)
        out["img_no_norm"] = denorm_image(imgs, norm_type)
        # which fields exist follows the model's scene representation (reference inference.py:350-379)
        if "pts3d_cam" in out:
            out["depth_z"] = out["pts3d_cam"][..., 2:3]
        if "ray_directions" in out:
            out["intrinsics"] = intrinsics_from_rays(out["ray_directions"])
        if "cam_trans" in out and "cam_quats" in out:
            out["camera_poses"] = pose_matrices(out["cam_quats"], out["cam_trans"])
        if apply_mask:
            final = out.get("non_ambiguous_mask")
            if apply_confidence_mask and "conf" in out:
                cm = quantile_mask(out["conf"], confidence_percentile / 100.0)
                final = cm if final is None else mask_and(final, cm)
            if mask_edges and final is not None and "pts3d" in out:
                if "pts3d_cam" not in out:   # the reference reads processed_output["depth_z"] here (inference.py:440)
                    raise KeyError("depth_z")
                final = edge_mask(out["pts3d"], out["pts3d_cam"], final, edge_normal_threshold, edge_depth_threshold)
            if final is not None:
                _apply_final_mask(out, final.contiguous())
        scale = out.pop("metric_scaling_factor")
        res = []
        for i in range(V):
            d = {k: v[i:i + 1] for k, v in out.items()}
            d["metric_scaling_factor"] = scale
            res.append(d)
        return res


def _apply_final_mask(out: Dict[str, torch.Tensor], m: torch.Tensor) -> None:
    """Zero the dense geometry outside the mask and attach it (reference inference.py:452-476); absent keys are skipped."""
    pts_cam = out.get("pts3d_cam")
    out["pts3d"] = _masked(out["pts3d"], m, 3)
    if pts_cam is not None:
        out["depth_z"] = _masked(pts_cam, m, 1, in_stride=3, in_offset=2)
        out["pts3d_cam"] = _masked(pts_cam, m, 3)
    if "depth_along_ray" in out:
        out["depth_along_ray"] = _masked(out["depth_along_ray"], m, 1)
    out["mask"] = m.unsqueeze(-1)


def _masked(x, m, width, in_stride=None, in_offset=0):
    return mask_dense(x, m, width, in_stride, in_offset)
