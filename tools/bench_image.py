"""Micro-benchmark of the input side (load_images: resize + crop + normalise), run under gpurun:

    python tools/bench_image.py [W H] [frames]

Synthetic uint8 frames resident in HBM (a pool larger than the 126 MB L2) -> mapanything_b200.image.resize_crop_normalize
(the two integer kernels of csrc/image.cu), CUDA events; algorithmic bytes per frame = source rows read + 8-bit
intermediate written and read + fp32 planes written.  Beside it: the reference's host pipeline for the same frame
(PIL resize + crop + the torchvision arithmetic), one host thread like the reference's loop, and the end-to-end
host frame -> device tensor path (pinned upload + kernels).  One JSON line."""
import json
import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "map-anything_b200"))
from mapanything_b200.image import _Uploader, find_closest_aspect_ratio, resize_crop_normalize, resize_plan  # noqa: E402


def main():
    W, H = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (1920, 1080)
    n = int(sys.argv[3]) if len(sys.argv) > 3 else 64
    target = find_closest_aspect_ratio(W / H, 518)
    rng = np.random.default_rng(0)
    host = [rng.integers(0, 256, (H, W, 3), dtype=np.uint8) for _ in range(4)]
    pool = torch.stack([torch.from_numpy(host[i % 4]).cuda().roll(i, 1) for i in range(n)])  # n * W*H*3 bytes, > L2
    for i in range(3):
        resize_crop_normalize(pool, target)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for rep in range(4):
        resize_crop_normalize(pool, target)   # one batched launch of each kernel for the n frames
    e.record()
    torch.cuda.synchronize()
    ms = s.elapsed_time(e) / (4 * n)
    s.record()
    for f in pool[:16]:
        resize_crop_normalize(f, target)      # frame by frame: launch-latency bound
    e.record()
    torch.cuda.synchronize()
    ms_single = s.elapsed_time(e) / 16
    rw, rh, filt, left, top = resize_plan(W, H, target)
    tw, th = target
    # rows of the source the vertical windows touch ~ all H rows; columns read ~ W
    bytes_alg = H * W * 3 + 2 * H * tw * 3 + th * tw * 3 * 4
    # reference host pipeline on one frame
    import PIL.Image

    mean = torch.tensor((0.485, 0.456, 0.406)).view(3, 1, 1)
    std = torch.tensor((0.229, 0.224, 0.225)).view(3, 1, 1)

    def normalize_u8(a, _norm):  # torchvision ToTensor + Normalize, as the reference applies them (image.py:291-296)
        return torch.from_numpy(a.copy()).permute(2, 0, 1).contiguous().float().div(255).sub_(mean).div_(std)[None]

    t0 = time.perf_counter()
    reps = 5
    for i in range(reps):
        im = PIL.Image.fromarray(host[i % 4]).resize((rw, rh), resample=filt).crop((left, top, left + tw, top + th))
        normalize_u8(np.asarray(im), "dinov2")
    host_ms = (time.perf_counter() - t0) / reps * 1e3
    # host frame -> device tensor, end to end (pinned staging upload + kernels)
    up = _Uploader(torch.device("cuda", 0))
    for i in range(4):
        resize_crop_normalize(up.upload(host[i]), target)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(32):
        resize_crop_normalize(up.upload(host[i % 4]), target)
    torch.cuda.synchronize()
    e2e_ms = (time.perf_counter() - t0) / 32 * 1e3
    print(json.dumps({
        "workload": f"{W}x{H} RGB frame -> {tw}x{th} fp32 normalised (resized {rw}x{rh}, filter {'LANCZOS' if filt == 1 else 'BICUBIC'})",
        "gpu_ms_per_frame": round(ms, 4), "gpu_frames_per_s": round(1e3 / ms, 1), "batch": n,
        "gpu_ms_per_frame_unbatched": round(ms_single, 4),
        "algorithmic_MB_per_frame": round(bytes_alg / 1e6, 2), "achieved_GBps": round(bytes_alg / ms / 1e6, 1),
        "hbm_peak_GBps": 6542.7, "frac": round(bytes_alg / ms / 1e6 / 6542.7, 3),
        "host_reference_ms_per_frame": round(host_ms, 2), "host_to_device_e2e_ms_per_frame": round(e2e_ms, 3),
        "pool_MB": round(n * W * H * 3 / 1e6, 1)}))


if __name__ == "__main__":
    main()
