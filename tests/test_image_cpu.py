"""Input side (load_images): the numpy restatement of the reference path (oracle/image.py) against
  * tests/golden/image.npz -- outputs of the reference's own load_images (oracle/make_golden.py), and
  * the installed Pillow (the third-party library whose 8-bit resampling the reference calls), bit for bit;
and the HOST entry point ma_resample_coeffs of the C-ABI library against the same coefficient tables.  No GPU needed."""
import ctypes as C
from pathlib import Path

import numpy as np
import pytest
import torch

from oracle import image as OI

GOLD = Path(__file__).parent / "golden" / "image.npz"
CASES = {
    "fixed": dict(resize_mode="fixed_size", size=(140, 98)),
    "square": dict(resize_mode="square", size=112),
    "longest": dict(resize_mode="longest_side", size=154),
    "mapping": dict(resize_mode="fixed_mapping"),
}


def write_case(gold, name, folder):
    import PIL.Image

    i = 0
    while f"{name}_src{i}" in gold:
        PIL.Image.fromarray(gold[f"{name}_src{i}"]).save(folder / f"img_{i:02d}.png")
        i += 1
    (folder / "notes.txt").write_text("not an image")
    return i


def check_view(gold, name, i, view):
    img = view["img"].cpu().numpy()
    if f"{name}_img{i}" in gold:
        assert np.array_equal(img, gold[f"{name}_img{i}"])
    else:
        st = int(gold[f"{name}_img{i}_stride"])
        assert tuple(img.shape) == tuple(gold[f"{name}_img{i}_shape"])
        assert np.array_equal(img[:, :, ::st, ::st], gold[f"{name}_img{i}_sample"])
        assert np.array_equal(img.astype(np.float64).sum(axis=(0, 2, 3)), gold[f"{name}_img{i}_sum"])
    assert (view["true_shape"] == gold[f"{name}_true_shape{i}"]).all()
    assert view["idx"] == i and view["instance"] == str(i) and view["data_norm_type"] == ["dinov2"]


@pytest.mark.parametrize("name", list(CASES))
def test_oracle_load_images_matches_reference_golden(name, tmp_path):
    gold = np.load(GOLD)
    n = write_case(gold, name, tmp_path)
    views = OI.load_images(str(tmp_path), **CASES[name])
    assert len(views) == n
    for i, v in enumerate(views):
        assert v["img"].dtype == torch.float32
        check_view(gold, name, i, v)


@pytest.mark.parametrize("W,H,ow,oh,filt", [
    (640, 480, 518, 389, OI.LANCZOS), (97, 131, 140, 189, OI.BICUBIC), (1000, 333, 518, 173, OI.LANCZOS),
    (300, 300, 518, 518, OI.BICUBIC), (640, 480, 640, 300, OI.LANCZOS), (37, 53, 37, 80, OI.BICUBIC), (5, 4, 14, 14, OI.BICUBIC),
])
def test_oracle_resize_matches_pillow(W, H, ow, oh, filt):
    import PIL.Image

    img = np.random.default_rng(W + H).integers(0, 256, (H, W, 3), dtype=np.uint8)
    ref = np.asarray(PIL.Image.fromarray(img).resize((ow, oh), resample=filt))
    assert np.array_equal(OI.pil_resize_u8(img, (ow, oh), filt), ref)


def test_resize_plan_and_target_sizes():
    assert OI.find_closest_aspect_ratio(4 / 3, 518) == (518, 392)
    assert OI.find_closest_aspect_ratio(0.5, 518) == (252, 518)
    assert OI.find_closest_aspect_ratio(16 / 9, 512) == (512, 288)
    assert OI.target_size_for([1.5, 1.3], "longest_side", 518, 14, 518) == (518, 364)
    assert OI.target_size_for([0.75], "longest_side", 518, 14, 518) == (392, 518)
    rw, rh, filt, left, top = OI.resize_plan(1920, 1080, (518, 294))
    assert (rw, rh, filt) == (522, 294, OI.LANCZOS) and (left, top) == (2, 0)
    rw, rh, filt, left, top = OI.resize_plan(100, 80, (140, 140))
    assert filt == OI.BICUBIC and rh == 140 and rw == 175 and (left, top) == (17, 0)


def _lib():
    import sys

    sys.path.insert(0, str(Path(__file__).resolve().parents[1] / "map-anything_b200"))
    from mapanything_b200 import _lib

    return _lib.load()


@pytest.mark.parametrize("i,o,filt", [(1920, 921, 1), (480, 389, 1), (97, 140, 3), (300, 518, 3), (64, 64, 1), (7, 3, 1)])
def test_cabi_resample_coeffs_host(i, o, filt):
    """ma_resample_coeffs is a host function of the shared library: same windows and 22-bit coefficients as the oracle."""
    lib = _lib()
    ks = C.c_int(0)
    assert lib.ma_resample_coeffs(i, o, filt, C.byref(ks), None, None) == 0
    bounds = np.zeros((o, 2), np.int32)
    coeffs = np.zeros((ks.value, o), np.int32)
    assert lib.ma_resample_coeffs(i, o, filt, C.byref(ks), bounds.ctypes.data, coeffs.ctypes.data) == 0
    if i == o:
        assert ks.value == 1 and (bounds[:, 0] == np.arange(o)).all() and (coeffs == 1 << 22).all()
        return
    b2, k2 = OI.resample_coeffs(i, o, filt)
    assert ks.value == k2.shape[1]
    assert np.array_equal(bounds, b2) and np.array_equal(coeffs.T, k2)
    assert lib.ma_resample_coeffs(0, o, filt, C.byref(ks), None, None) != 0  # bad size -> status + message
    assert b"bad sizes" in lib.ma_last_error()


def test_load_images_validation_errors_need_no_gpu():
    """Argument validation of the drop-in load_images happens before any device work, with the reference's messages."""
    import sys

    sys.path.insert(0, str(Path(__file__).resolve().parents[1] / "map-anything_b200"))
    from mapanything_b200.image import load_images

    with pytest.raises(ValueError, match="Resize_mode must be one of"):
        load_images([], resize_mode="bogus")
    with pytest.raises(ValueError, match="Size parameter is required"):
        load_images([], resize_mode="square")
    with pytest.raises(ValueError, match="Size must be an int"):
        load_images([], resize_mode="longest_side", size=(1, 2))
    with pytest.raises(ValueError, match=r"tuple/list of \(width, height\)"):
        load_images([], resize_mode="fixed_size", size=5)
    with pytest.raises(ValueError, match="Bad folder_or_list"):
        load_images(42)
    with pytest.raises(ValueError, match="Unknown image normalization type"):
        load_images([], norm_type="nope")


# ------------------------------------------------------------------------------------------------ preprocess_inputs
PGOLD = Path(__file__).parent / "golden" / "preprocess_inputs.npz"


def check_preprocessed(gold, name, views, rays_tol=2e-3, source_views=None):
    """views: output of a preprocess_inputs implementation on oracle.make_golden.synth_multimodal_views()."""
    assert len(views) == 5
    for i, v in enumerate(views):
        assert v["data_norm_type"] == ["dinov2"]
        img = v["img"].cpu().numpy()
        if f"{name}_v{i}_img" in gold:
            assert np.array_equal(img, gold[f"{name}_v{i}_img"])
        else:
            assert tuple(img.shape) == tuple(gold[f"{name}_v{i}_img_shape"])
            assert np.array_equal(img[:, :, ::3, ::3], gold[f"{name}_v{i}_img_sample"])
            assert np.array_equal(img.astype(np.float64).sum(axis=(0, 2, 3)), gold[f"{name}_v{i}_img_sum"])
        for key in ("depth_z", "intrinsics", "camera_poses", "is_metric_scale"):
            gk = f"{name}_v{i}_{key}"
            if key == "camera_poses" and f"{name}_v{i}_pose_q" in gold:
                q, t = v["camera_poses"]
                assert np.array_equal(q.cpu().numpy(), gold[f"{name}_v{i}_pose_q"])
                assert np.array_equal(t.cpu().numpy(), gold[f"{name}_v{i}_pose_t"])
                continue
            if gk not in gold:
                assert key not in v, (i, key)
                continue
            got = v[key].cpu().numpy()
            assert got.shape == gold[gk].shape and got.dtype == gold[gk].dtype, (i, key, got.dtype, gold[gk].dtype)
            if key == "intrinsics" and i == 2:  # recovered from ray directions (least squares): tolerance, not bits
                assert np.abs(got - gold[gk]).max() <= rays_tol * np.abs(gold[gk]).max()
            else:
                assert np.array_equal(got, gold[gk]), (i, key)
        assert "ray_directions" not in v
    assert views[3]["instance"] == "x"


@pytest.mark.parametrize("name", ["fixed", "square", "longest"])
def test_oracle_preprocess_inputs_matches_reference_golden(name):
    from oracle.make_golden import PREPROCESS_CASES, synth_multimodal_views

    gold = np.load(PGOLD)
    check_preprocessed(gold, name, OI.preprocess_inputs(synth_multimodal_views(), **PREPROCESS_CASES[name]))


def test_oracle_nearest_matches_opencv():
    import cv2

    rng = np.random.default_rng(0)
    for _ in range(40):
        sh, sw, dh, dw = (int(v) for v in rng.integers(5, 400, 4))
        a = rng.random((sh, sw), dtype=np.float32)
        ref = cv2.resize(a, (dw, dh), interpolation=cv2.INTER_NEAREST)
        assert np.array_equal(a[OI.nearest_indices(sh, dh)][:, OI.nearest_indices(sw, dw)], ref)


def test_preprocess_inputs_validation_errors_need_no_gpu():
    from mapanything_b200.image import preprocess_inputs

    with pytest.raises(ValueError, match="input_views cannot be empty"):
        preprocess_inputs([])
    with pytest.raises(ValueError, match="No valid images found"):
        preprocess_inputs([{"depth_z": np.zeros((4, 4), np.float32)}])
    with pytest.raises(ValueError, match=r"Expected array shape \(H, W, 3\)"):
        preprocess_inputs([{"img": np.zeros((4, 4), np.uint8)}])
    with pytest.raises(ValueError, match="Unsupported image type"):
        preprocess_inputs([{"img": "file.png"}])


@pytest.mark.parametrize("i,o,filt", [(1920, 522, 1), (4032, 522, 1), (97, 140, 3), (64, 64, 1)])
def test_cabi_pack_coeffs_host(i, o, filt):
    """ma_resample_pack_coeffs (host): the byte planes recombine to the 22-bit coefficients; 4 taps per word, zero padded."""
    lib = _lib()
    ks = C.c_int(0)
    assert lib.ma_resample_coeffs(i, o, filt, C.byref(ks), None, None) == 0
    bounds = np.zeros((o, 2), np.int32)
    coeffs = np.zeros((ks.value, o), np.int32)
    assert lib.ma_resample_coeffs(i, o, filt, C.byref(ks), bounds.ctypes.data, coeffs.ctypes.data) == 0
    nw = (ks.value + 3) // 4
    packed = np.zeros((3, nw, o), np.uint32)
    assert lib.ma_resample_pack_coeffs(coeffs.ctypes.data, ks.value, o, packed.ctypes.data) == 0
    b = packed.view(np.uint8).reshape(3, nw, o, 4)                      # little endian: byte b of a word = tap 4j + b
    taps = b.transpose(0, 1, 3, 2).reshape(3, nw * 4, o).astype(np.int64)
    k = taps[0] + 256 * taps[1] + 65536 * taps[2].astype(np.uint8).view(np.int8).astype(np.int64).reshape(taps[2].shape)
    assert np.array_equal(k[:ks.value], coeffs) and not k[ks.value:].any()
    big = np.full((1, 1), 1 << 24, np.int32)
    assert lib.ma_resample_pack_coeffs(big.ctypes.data, 1, 1, packed.ctypes.data) != 0


def test_host_planning_matches_oracle_on_random_sizes():
    """Host-side planning of the drop-in (target size, resize plan, crop box, nearest-neighbour offsets, intrinsics update)
    against the oracle restatement of the reference over many random shapes -- no GPU involved."""
    from mapanything_b200 import image as PI

    rng = np.random.default_rng(42)
    for _ in range(400):
        W1, H1 = int(rng.integers(16, 5000)), int(rng.integers(16, 4000))
        mode = ["fixed_mapping", "square", "longest_side", "fixed_size"][int(rng.integers(0, 4))]
        size = None if mode == "fixed_mapping" else (int(rng.integers(28, 800)) if mode != "fixed_size"
                                                    else (int(rng.integers(28, 800)), int(rng.integers(28, 800))))
        ratios = [W1 / H1, float(rng.uniform(0.4, 2.5))]
        rset = int(rng.choice([518, 512]))
        target = PI._target_size(ratios, mode, size, 14, rset, False)
        assert tuple(target) == tuple(OI.target_size_for(ratios, mode, size, 14, rset))
        if min(target) < 14:
            continue
        assert PI.resize_plan(W1, H1, target) == OI.resize_plan(W1, H1, target)
        rw, rh, _, _, _ = PI.resize_plan(W1, H1, target)
        assert np.array_equal(PI._nearest_indices(W1, rw), OI.nearest_indices(W1, rw))
        for dtype in (np.float32, np.float64):
            K = np.array([[rng.uniform(0.5, 2) * W1, 0, W1 / 2 + rng.uniform(-20, 20)],
                          [0, rng.uniform(0.5, 2) * W1, H1 / 2 + rng.uniform(-20, 20)], [0, 0, 1]], dtype)
            scale = max(np.array(target) / np.array((W1, H1))) + 1e-8
            a = PI._camera_matrix_of_crop(K, np.array((W1, H1)), np.array((rw, rh)), scaling=scale)
            b = OI.camera_matrix_of_crop(K, np.array((W1, H1)), np.array((rw, rh)), scaling=scale)
            assert a.dtype == b.dtype == dtype and np.array_equal(a, b)
            a2 = PI._camera_matrix_of_crop(a, (rw, rh), target, offset_factor=0.5)
            b2 = OI.camera_matrix_of_crop(b, (rw, rh), target, offset_factor=0.5)
            assert np.array_equal(a2, b2)


@pytest.mark.parametrize("i,o,filt", [(1920, 522, 1), (4032, 522, 1), (8000, 130, 1), (97, 140, 3), (300, 518, 3)])
def test_dp4a_decomposition_is_exact_in_int32(i, o, filt):
    """The dp4a kernel accumulates three byte-plane partial sums in int32 and recombines them as s0 + 256 s1 + 65536 s2.
    Emulated here with numpy int32 (wrap-around) arithmetic for worst-case (all 255) and random pixels: identical to the
    direct 64-bit sum, which itself stays inside int32 (as it must for Pillow's own int accumulator)."""
    b, k = OI.resample_coeffs(i, o, filt)          # [out][ksize]
    k = k.astype(np.int64)
    k0, k1, k2 = k & 255, (k >> 8) & 255, k >> 16
    assert np.array_equal(k0 + 256 * k1 + 65536 * k2, k) and k2.min() >= -128 and k2.max() <= 127
    rng = np.random.default_rng(i + o)
    for px in (np.full(k.shape, 255, np.int64), rng.integers(0, 256, k.shape).astype(np.int64),
               np.where(k > 0, 255, 0).astype(np.int64), np.where(k < 0, 255, 0).astype(np.int64)):
        direct = (px * k).sum(1) + (1 << 21)
        assert direct.max() < 2 ** 31 and direct.min() >= -2 ** 31
        with np.errstate(over="ignore"):
            s0 = (px * k0).sum(1).astype(np.int32)
            s1 = (px * k1).sum(1).astype(np.int32)
            s2 = (px * k2).sum(1).astype(np.int32)
            acc = np.int32(1 << 21) + s0 + s1 * np.int32(256) + s2 * np.int32(65536)
        assert np.array_equal(acc.astype(np.int64), direct)
