"""Geometry helpers the reference's demos call on infer()'s outputs (mapanything/utils/geometry.py), on the GPU."""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from . import _lib, ops
from ._lib import check


def _f32(t: torch.Tensor) -> torch.Tensor:
    if not t.is_cuda:
        raise ValueError("mapanything_b200.geometry runs on the GPU: pass CUDA tensors (infer() returns them)")
    return t.contiguous().float()


def depthmap_to_world_frame(depthmap: torch.Tensor, intrinsics: torch.Tensor, camera_pose: Optional[torch.Tensor] = None
                            ) -> Tuple[torch.Tensor, torch.Tensor]:
    """Depth image -> point map in the world frame + mask of valid (non-zero depth) pixels; same contract as the
    reference (geometry.py:76-114): depthmap HxW or BxHxW, intrinsics 3x3 or Bx3x3, camera_pose 4x4 or Bx4x4 or None."""
    squeeze = depthmap.dim() == 2
    d = _f32(depthmap[None] if squeeze else depthmap)
    K = _f32(intrinsics[None] if intrinsics.dim() == 2 else intrinsics)
    P = None
    if camera_pose is not None:
        P = _f32(camera_pose[None] if camera_pose.dim() == 2 else camera_pose)
    n, H, W = d.shape
    if K.shape != (n, 3, 3) or (P is not None and P.shape != (n, 4, 4)):
        raise ValueError(f"depthmap {tuple(d.shape)}, intrinsics {tuple(K.shape)} and camera_pose batch sizes do not match")
    pts = torch.empty(n, H, W, 3, device=d.device, dtype=torch.float32)
    valid = torch.empty(n, H, W, device=d.device, dtype=torch.bool)
    check(_lib.load().ma_depthmap_to_world(d.data_ptr(), K.data_ptr(), None if P is None else P.data_ptr(), pts.data_ptr(),
                                           valid.data_ptr(), n, H, W, torch.cuda.current_stream().cuda_stream),
          "ma_depthmap_to_world")
    ops._count()
    return (pts[0], valid[0]) if squeeze else (pts, valid)


def depthmap_to_camera_frame(depthmap: torch.Tensor, intrinsics: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """geometry.py:18-73."""
    return depthmap_to_world_frame(depthmap, intrinsics, None)
