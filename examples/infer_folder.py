#!/usr/bin/env python
"""Images-only inference on a folder of frames with the B200 drop-in -- the flow of the reference's
scripts/demo_images_only_inference.py (load_images -> model.infer -> depthmap_to_world_frame), every step on the GPU:

    python examples/infer_folder.py /path/to/frames --checkpoint model.safetensors --out scene.npz [--ply scene.ply]

Without --checkpoint the model runs on random-init weights (shapes / speed only).  Writes per-view depth, intrinsics,
camera poses, confidence and masks, plus the fused point cloud of all valid pixels."""
import argparse
import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "map-anything_b200"))
from mapanything_b200 import MapAnything, load_images, mapanything_config  # noqa: E402
from mapanything_b200.geometry import depthmap_to_world_frame  # noqa: E402


def load_weights(model, path):
    if path.endswith(".safetensors"):
        from safetensors.torch import load_file

        state = load_file(path)
    else:
        state = torch.load(path, map_location="cpu", weights_only=False)
        state = state.get("model", state)
    print(model.load_state_dict(state, strict=False))


def write_ply(path, pts, rgb):
    with open(path, "wb") as f:
        f.write((f"ply\nformat binary_little_endian 1.0\nelement vertex {len(pts)}\nproperty float x\nproperty float y\n"
                 "property float z\nproperty uchar red\nproperty uchar green\nproperty uchar blue\nend_header\n").encode())
        rec = np.empty(len(pts), dtype=[("p", "<f4", 3), ("c", "u1", 3)])
        rec["p"], rec["c"] = pts, rgb
        f.write(rec.tobytes())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("folder")
    ap.add_argument("--checkpoint", default=None)
    ap.add_argument("--out", default="scene.npz")
    ap.add_argument("--ply", default=None)
    ap.add_argument("--stride", type=int, default=1)
    ap.add_argument("--memory-efficient-inference", action="store_true")
    args = ap.parse_args()

    model = MapAnything(**mapanything_config())
    if args.checkpoint:
        load_weights(model, args.checkpoint)
    model = model.to("cuda").eval()

    t0 = time.perf_counter()
    views = load_images(args.folder, stride=args.stride)            # decode on the host, resize / crop / normalise on the GPU
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    preds = model.infer(views, memory_efficient_inference=args.memory_efficient_inference)
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(f"{len(views)} views: load_images {t1 - t0:.2f} s, infer {t2 - t1:.3f} s ({len(views) / (t2 - t1):.1f} views/s)")

    out, cloud, colors = {}, [], []
    for i, p in enumerate(preds):
        depth = p["depth_z"][0].squeeze(-1)
        pts, valid = depthmap_to_world_frame(depth, p["intrinsics"][0], p["camera_poses"][0])
        mask = p["mask"][0].squeeze(-1) & valid
        out[f"depth_{i}"] = depth.cpu().numpy()
        out[f"intrinsics_{i}"] = p["intrinsics"][0].cpu().numpy()
        out[f"pose_{i}"] = p["camera_poses"][0].cpu().numpy()
        out[f"conf_{i}"] = p["conf"][0].cpu().numpy()
        out[f"mask_{i}"] = mask.cpu().numpy()
        cloud.append(pts[mask].cpu().numpy())
        colors.append((p["img_no_norm"][0][mask] * 255).clamp(0, 255).byte().cpu().numpy())
    out["points"], out["colors"] = np.concatenate(cloud), np.concatenate(colors)
    np.savez_compressed(args.out, **out)
    print(f"wrote {args.out}: {len(out['points'])} points")
    if args.ply:
        write_ply(args.ply, out["points"], out["colors"])
        print(f"wrote {args.ply}")


if __name__ == "__main__":
    main()
