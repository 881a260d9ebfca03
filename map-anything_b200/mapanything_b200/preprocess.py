"""preprocess_input_views_for_inference (reference mapanything/utils/inference.py:202-291): converts the user-facing
view keys into the model's internal keys.  Per-pixel conversions run in the ma_* preprocessing kernels."""
from __future__ import annotations

from typing import Any, Dict, List

import torch


def preprocess_input_views_for_inference(views: List[Dict[str, Any]]) -> List[Dict[str, Any]]:
    processed_views = []
    for view_idx, view in enumerate(views):
        pv = dict(view)  # shallow copy: never modify the caller's dict beyond the device move done by infer()
        if "intrinsics" in view or "ray_directions" in view or "depth_z" in view or "camera_poses" in view:
            from . import geometric_inputs  # kernels for the multi-modal path

            pv = geometric_inputs.convert_view(pv, view, view_idx)
        if "is_metric_scale" not in pv:
            batch_size = view["img"].shape[0]
            pv["is_metric_scale"] = torch.ones(batch_size, dtype=torch.bool, device=view["img"].device)
        if "ray_directions" in pv:
            pv["ray_directions_cam"] = pv.pop("ray_directions")
        processed_views.append(pv)
    return processed_views
