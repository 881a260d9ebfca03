"""Input side of the drop-in: `load_images` with the reference's signature and results (mapanything/utils/image.py:134-332),
resize / crop / normalise on the GPU.

The reference decodes every file with Pillow, resizes it with PIL.Image.resize (LANCZOS when shrinking, BICUBIC when
enlarging; cropping.py:188-275), centre-crops (cropping.py:385-467) and normalises with torchvision (image.py:291-296),
all on the host.  Here the file is still DECODED on the host (Pillow), then the raw RGB bytes go to the device once and
two integer kernels (csrc/image.cu) reproduce Pillow's 8-bit resampling bit for bit, fused with the crop and the
normalisation: the returned views hold `img` as a CUDA fp32 (1, 3, H, W) tensor that `MapAnything.infer` consumes without
another host->device copy.  There is no host fallback for the resize.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib, ops
from ._lib import check
from .inference import IMAGE_NORMALIZATION_DICT

try:  # same optional dependency as the reference (image.py:24-32)
    from pillow_heif import register_heif_opener

    register_heif_opener()
    heif_support_enabled = True
except ImportError:
    heif_support_enabled = False

MA_FILTER_LANCZOS, MA_FILTER_BICUBIC = 1, 3

# Fixed aspect-ratio -> (W, H) tables of the released model (image.py:40-65)
RESOLUTION_MAPPINGS = {
    518: {
        1.000: (518, 518), 1.321: (518, 392), 1.542: (518, 336), 1.762: (518, 294), 2.056: (518, 252), 3.083: (518, 168),
        0.757: (392, 518), 0.649: (336, 518), 0.567: (294, 518), 0.486: (252, 518),
    },
    512: {
        1.000: (512, 512), 1.333: (512, 384), 1.524: (512, 336), 1.778: (512, 288), 2.000: (512, 256), 3.200: (512, 160),
        0.750: (384, 512), 0.656: (336, 512), 0.562: (288, 512), 0.500: (256, 512),
    },
}
ASPECT_RATIO_KEYS = {k: sorted(v.keys()) for k, v in RESOLUTION_MAPPINGS.items()}


def find_closest_aspect_ratio(aspect_ratio: float, resolution_set: int) -> Tuple[int, int]:
    """(target_width, target_height) of the table entry whose aspect-ratio key is closest (image.py:74-90)."""
    closest_key = min(ASPECT_RATIO_KEYS[resolution_set], key=lambda x: abs(x - aspect_ratio))
    return RESOLUTION_MAPPINGS[resolution_set][closest_key]


# ------------------------------------------------------------------------------------------------ device resampler
class _AxisTables:
    """Pillow's resampling windows of one axis (in_size -> out_size): host bounds for planning, device copies for the kernels."""

    __slots__ = ("bounds", "d_bounds", "d_coeffs", "out_size")

    def __init__(self, in_size: int, out_size: int, filt: int, device: torch.device):
        lib = _lib.load()
        ks = C.c_int(0)
        check(lib.ma_resample_coeffs(in_size, out_size, filt, C.byref(ks), None, None), "ma_resample_coeffs")
        self.bounds = np.empty((out_size, 2), np.int32)
        coeffs = np.empty((ks.value, out_size), np.int32)
        check(lib.ma_resample_coeffs(in_size, out_size, filt, C.byref(ks), self.bounds.ctypes.data, coeffs.ctypes.data),
              "ma_resample_coeffs")
        self.d_bounds = torch.from_numpy(self.bounds).to(device)
        self.d_coeffs = torch.from_numpy(coeffs).to(device)
        self.out_size = out_size


_TABLES: Dict[tuple, _AxisTables] = {}


def _tables(in_size: int, out_size: int, filt: int, device: torch.device) -> _AxisTables:
    key = (in_size, out_size, filt, str(device))
    t = _TABLES.get(key)
    if t is None:
        if len(_TABLES) > 64:
            _TABLES.clear()
        t = _TABLES[key] = _AxisTables(in_size, out_size, filt, device)
    return t


def resize_plan(W1: int, H1: int, target: Tuple[int, int]) -> Tuple[int, int, int, int, int]:
    """(resized W, resized H, filter, crop left, crop top) the reference applies to a W1 x H1 image for target (W, H):
    scale so that the resized image contains the crop (cropping.py:236-247), centred crop (cropping.py:440-446)."""
    out_res = np.array(target)
    in_res = np.array((W1, H1))
    scale_final = max(out_res / in_res) + 1e-8
    rw, rh = (int(v) for v in np.floor(in_res * scale_final).astype(int))
    filt = MA_FILTER_LANCZOS if scale_final < 1 else MA_FILTER_BICUBIC
    return rw, rh, filt, (rw - target[0]) // 2, (rh - target[1]) // 2


def resize_crop_normalize(img_u8: torch.Tensor, target: Tuple[int, int], norm_type: Optional[str] = "dinov2",
                          return_u8: bool = False):
    """img_u8: CUDA uint8 (H, W, 3), rows contiguous -> fp32 (1, 3, th, tw) normalised image (and / or the uint8
    (th, tw, 3) image when return_u8), bit-exact with PIL resize + crop + torchvision ToTensor/Normalize."""
    if not img_u8.is_cuda or img_u8.dtype != torch.uint8 or img_u8.dim() != 3 or img_u8.shape[2] != 3 or img_u8.stride(2) != 1 \
            or img_u8.stride(1) != 3:
        raise ValueError("resize_crop_normalize expects a CUDA uint8 (H, W, 3) tensor with packed RGB pixels")
    H1, W1, _ = img_u8.shape
    tw, th = int(target[0]), int(target[1])
    rw, rh, filt, left, top = resize_plan(W1, H1, (tw, th))
    dev = img_u8.device
    th_, tv_ = _tables(W1, rw, filt, dev), _tables(H1, rh, filt, dev)
    y0 = int(tv_.bounds[top, 0])
    y1 = int(tv_.bounds[top + th - 1, 0] + tv_.bounds[top + th - 1, 1])
    sx0 = int(th_.bounds[left, 0])
    sx1 = int(th_.bounds[left + tw - 1, 0] + th_.bounds[left + tw - 1, 1])
    lib = _lib.load()
    stream = torch.cuda.current_stream().cuda_stream
    tmp = torch.empty(y1 - y0, tw, 3, device=dev, dtype=torch.uint8)
    check(lib.ma_resample_h_u8rgb(img_u8.data_ptr(), img_u8.stride(0), y0, y1 - y0, sx0, sx1, th_.d_bounds.data_ptr(),
                                  th_.d_coeffs.data_ptr(), rw, left, tw, tmp.data_ptr(), stream), "ma_resample_h_u8rgb")
    out = out8 = None
    m3 = s3 = None
    if norm_type is not None:
        if norm_type not in IMAGE_NORMALIZATION_DICT:
            raise ValueError(
                f"Unknown image normalization type: {norm_type}. Available options: {list(IMAGE_NORMALIZATION_DICT.keys())}"
            )
        mean, std = IMAGE_NORMALIZATION_DICT[norm_type]
        m3, s3 = (C.c_float * 3)(*mean), (C.c_float * 3)(*std)
        out = torch.empty(1, 3, th, tw, device=dev, dtype=torch.float32)
    if return_u8 or out is None:
        out8 = torch.empty(th, tw, 3, device=dev, dtype=torch.uint8)
    check(lib.ma_resample_v_norm_u8rgb(tmp.data_ptr(), tw, y0, tv_.d_bounds.data_ptr(), tv_.d_coeffs.data_ptr(), rh, top, th,
                                       m3, s3, None if out is None else out.data_ptr(),
                                       None if out8 is None else out8.data_ptr(), stream), "ma_resample_v_norm_u8rgb")
    ops._count(2)
    if out is None:
        return out8
    return (out, out8) if return_u8 else out


class _Uploader:
    """Host -> device copies of decoded frames through two reusable pinned staging buffers."""

    def __init__(self, device: torch.device):
        self.device = device
        self.bufs: List[Optional[torch.Tensor]] = [None, None]
        self.events: List[Optional[torch.cuda.Event]] = [None, None]
        self.i = 0

    def upload(self, arr: np.ndarray) -> torch.Tensor:
        arr = np.ascontiguousarray(arr)
        n = arr.size
        j = self.i
        self.i ^= 1
        if self.events[j] is not None:
            self.events[j].synchronize()  # the previous copy out of this staging buffer has completed
        if self.bufs[j] is None or self.bufs[j].numel() < n:
            self.bufs[j] = torch.empty(max(n, 1 << 20), dtype=torch.uint8, pin_memory=True)
        self.bufs[j][:n].numpy()[...] = arr.reshape(-1)
        dst = torch.empty(arr.shape, dtype=torch.uint8, device=self.device)
        dst.view(-1).copy_(self.bufs[j][:n], non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        self.events[j] = ev
        return dst


def _target_size(aspect_ratios: Sequence[float], resize_mode: str, size, patch_size: int, resolution_set: int, verbose: bool):
    """One (W, H) for all images from their average aspect ratio (image.py:240-289)."""
    average_aspect_ratio = sum(aspect_ratios) / len(aspect_ratios)
    if verbose:
        print(f"Calculated average aspect ratio: {average_aspect_ratio:.3f} from {len(aspect_ratios)} images")
    if resize_mode == "fixed_mapping":
        return tuple(find_closest_aspect_ratio(average_aspect_ratio, resolution_set))
    if resize_mode == "square":
        side = round((size // patch_size)) * patch_size
        return (side, side)
    if resize_mode == "longest_side":
        if average_aspect_ratio >= 1:  # width is the longest side
            return (size, round((size // patch_size) / average_aspect_ratio) * patch_size)
        return (round((size // patch_size) * average_aspect_ratio) * patch_size, size)
    return ((size[0] // patch_size) * patch_size, (size[1] // patch_size) * patch_size)  # fixed_size


def load_images(
    folder_or_list,
    resize_mode="fixed_mapping",
    size=None,
    norm_type="dinov2",
    patch_size=14,
    verbose=False,
    bayer_format=False,
    resolution_set=518,
    stride=1,
    device=None,
):
    """Same contract as the reference `load_images` (image.py:134-332): a list of view dicts with `img` (1, 3, H, W) fp32,
    `true_shape`, `idx`, `instance`, `data_norm_type`; every image is brought to ONE target size chosen from the average
    aspect ratio.  `img` lives on `device` (default: the current CUDA device)."""
    valid_resize_modes = ["fixed_mapping", "longest_side", "square", "fixed_size"]
    if resize_mode not in valid_resize_modes:
        raise ValueError(f"Resize_mode must be one of {valid_resize_modes}, got '{resize_mode}'")
    if resize_mode in ["longest_side", "square", "fixed_size"] and size is None:
        raise ValueError(f"Size parameter is required for resize_mode='{resize_mode}'")
    if resize_mode in ["longest_side", "square"]:
        if not isinstance(size, int):
            raise ValueError(f"Size must be an int for resize_mode='{resize_mode}', got {type(size)}")
    elif resize_mode == "fixed_size":
        if not isinstance(size, (tuple, list)) or len(size) != 2:
            raise ValueError(f"Size must be a tuple/list of (width, height) for resize_mode='fixed_size', got {size}")
        if not all(isinstance(x, int) for x in size):
            raise ValueError(f"Size values must be integers for resize_mode='fixed_size', got {size}")

    if isinstance(folder_or_list, str):
        if verbose:
            print(f"Loading images from {folder_or_list}")
        root, folder_content = folder_or_list, sorted(os.listdir(folder_or_list))
    elif isinstance(folder_or_list, list):
        if verbose:
            print(f"Loading a list of {len(folder_or_list)} images")
        root, folder_content = "", folder_or_list
    else:
        raise ValueError(f"Bad {folder_or_list=} ({type(folder_or_list)})")

    if norm_type not in IMAGE_NORMALIZATION_DICT:
        raise ValueError(
            f"Unknown image normalization type: {norm_type}. Available options: {list(IMAGE_NORMALIZATION_DICT.keys())}"
        )
    dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
    if dev.type != "cuda":
        raise RuntimeError("mapanything_b200.load_images resizes on the GPU: pass a CUDA device")

    import PIL.Image
    from PIL.ImageOps import exif_transpose

    exts = [".jpg", ".jpeg", ".png"] + ([".heic", ".heif"] if heif_support_enabled else [])
    exts = tuple(exts)
    # first pass: decode on the host (Pillow, like the reference), start the uploads, collect aspect ratios
    up = _Uploader(dev)
    loaded, aspect_ratios = [], []
    with torch.cuda.device(dev):
        for i, path in enumerate(folder_content):
            if i % stride != 0 or not path.lower().endswith(exts):
                continue
            try:
                if bayer_format:
                    import cv2

                    color_bayer = cv2.imread(os.path.join(root, path), cv2.IMREAD_UNCHANGED)
                    img = PIL.Image.fromarray(cv2.cvtColor(color_bayer, cv2.COLOR_BAYER_RG2BGR))
                    img = exif_transpose(img).convert("RGB")
                else:
                    img = exif_transpose(PIL.Image.open(os.path.join(root, path))).convert("RGB")
                W1, H1 = img.size
                loaded.append((path, up.upload(np.asarray(img)), W1, H1))
                aspect_ratios.append(W1 / H1)
            except Exception as e:  # noqa: BLE001  (the reference skips unreadable files the same way)
                if verbose:
                    print(f"Warning: Could not load {path}: {e}")
                continue
        if not loaded:
            raise ValueError("No valid images found")
        target_size = _target_size(aspect_ratios, resize_mode, size, patch_size, resolution_set, verbose)
        if verbose:
            print(f"Using target resolution {target_size[0]}x{target_size[1]} (W x H) for all images")
        # second pass: resize + crop + normalise on the device
        imgs = []
        for path, d_img, W1, H1 in loaded:
            t = resize_crop_normalize(d_img, target_size, norm_type)
            H2, W2 = t.shape[2], t.shape[3]
            if verbose:
                print(f" - Adding {path} with resolution {W1}x{H1} --> {W2}x{H2}")
            imgs.append(dict(img=t, true_shape=np.int32([[H2, W2]]), idx=len(imgs), instance=str(len(imgs)),
                             data_norm_type=[norm_type]))
    if verbose:
        print(f" (Found {len(imgs)} images)")
    return imgs
