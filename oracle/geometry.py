"""Oracle restatement of the geometry math on the MapAnything inference path (test infrastructure only).

Each function names the reference function it follows in /root/reference/mapanything/utils/geometry.py.
Pinned by tests/test_oracle_golden.py against fixtures produced by importing that file (oracle/make_golden.py).
"""
from __future__ import annotations

import numpy as np
import torch


# ------------------------------------------------------------------------------------------------
# rays <-> intrinsics
# ------------------------------------------------------------------------------------------------
def rays_from_intrinsics(intrinsics: torch.Tensor, height: int, width: int, unit_sphere: bool = True) -> torch.Tensor:
    """geometry.py:186-241 get_rays_in_camera_frame (directions only). intrinsics (B,3,3) -> (B,H,W,3)."""
    k = intrinsics if intrinsics.dim() == 3 else intrinsics[None]
    dev = k.device
    xs = torch.arange(width, device=dev).float()[None, None, :]
    ys = torch.arange(height, device=dev).float()[None, :, None]
    fx, fy = k[:, 0, 0].view(-1, 1, 1), k[:, 1, 1].view(-1, 1, 1)
    cx, cy = k[:, 0, 2].view(-1, 1, 1), k[:, 1, 2].view(-1, 1, 1)
    xx = ((xs - cx) / fx).expand(-1, height, width)
    yy = ((ys - cy) / fy).expand(-1, height, width)
    d = torch.stack([xx, yy, torch.ones_like(xx)], dim=-1)
    if unit_sphere:
        d = d / torch.norm(d, dim=-1, keepdim=True)
    return d if intrinsics.dim() == 3 else d[0]


def intrinsics_from_rays(rays: torch.Tensor) -> torch.Tensor:
    """geometry.py:304-447 recover_pinhole_intrinsics_from_ray_directions, default arguments.

    <= 1 MP: least squares of x = cx + fx*(dx/dz) (and y likewise) over a grid sampled every
    max(1, H//50) x max(1, W//50) pixels, solved through the 2x2 normal equations.
    > 1 MP: closed form from five key pixels."""
    squeeze = rays.dim() == 3
    if squeeze:
        rays = rays[None]
    b, h, w, _ = rays.shape
    dev = rays.device
    if h * w > 1_000_000:
        ch, cw = h // 2, w // 2
        qw, tqw, qh, tqh = w // 4, 3 * w // 4, h // 4, 3 * h // 4

        def unit_z(r):
            return r / r[:, 2:3]

        c = unit_z(rays[:, ch, cw].clone())
        left, right = unit_z(rays[:, ch, qw].clone()), unit_z(rays[:, ch, tqw].clone())
        top, bot = unit_z(rays[:, qh, cw].clone()), unit_z(rays[:, tqh, cw].clone())
        fx = ((qw - cw) / (left[:, 0] - c[:, 0]) + (tqw - cw) / (right[:, 0] - c[:, 0])) / 2
        cx = cw - fx * c[:, 0]
        fy = ((qh - ch) / (top[:, 1] - c[:, 1]) + (tqh - ch) / (bot[:, 1] - c[:, 1])) / 2
        cy = ch - fy * c[:, 1]
    else:
        hi = torch.arange(0, h, max(1, h // 50), device=dev)
        wi = torch.arange(0, w, max(1, w // 50), device=dev)
        samp = rays[:, hi[:, None], wi[None, :], :]
        xg = wi.float()[None, None, :].expand(b, hi.numel(), wi.numel()).reshape(b, -1)
        yg = hi.float()[None, :, None].expand(b, hi.numel(), wi.numel()).reshape(b, -1)
        rx = (samp[..., 0] / samp[..., 2]).reshape(b, -1)
        ry = (samp[..., 1] / samp[..., 2]).reshape(b, -1)

        def fit(ratio, target):
            a = torch.stack([torch.ones_like(ratio), ratio], dim=2)
            ata = torch.bmm(a.transpose(1, 2), a)
            atb = torch.bmm(a.transpose(1, 2), target.unsqueeze(2))
            sol = torch.linalg.solve(ata, atb).squeeze(2)
            return sol[:, 0], sol[:, 1]

        cx, fx = fit(rx, xg)
        cy, fy = fit(ry, yg)
    k = torch.zeros(b, 3, 3, device=dev)
    k[:, 0, 0], k[:, 1, 1], k[:, 0, 2], k[:, 1, 2], k[:, 2, 2] = fx, fy, cx, cy, 1.0
    return k[0] if squeeze else k


# ------------------------------------------------------------------------------------------------
# quaternions (x, y, z, w)
# ------------------------------------------------------------------------------------------------
def quat_to_rotmat(q: torch.Tensor) -> torch.Tensor:
    """geometry.py:601-652 quaternion_to_rotation_matrix (normalises first)."""
    squeeze = q.dim() == 1
    if squeeze:
        q = q[None]
    q = q / q.norm(dim=1, keepdim=True)
    x, y, z, w = q.unbind(1)
    r = torch.stack(
        [
            1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y),
            2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x),
            2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y),
        ],
        dim=1,
    ).view(-1, 3, 3)
    return r[0] if squeeze else r


def rotmat_to_quat(m: torch.Tensor) -> torch.Tensor:
    """geometry.py:655-713 rotation_matrix_to_quaternion (+ :716-742 helpers): best-conditioned of the four
    candidates, returned scalar-last with w >= 0."""
    batch = m.shape[:-2]
    m00, m01, m02, m10, m11, m12, m20, m21, m22 = m.reshape(batch + (9,)).unbind(-1)
    sq = torch.stack([1 + m00 + m11 + m22, 1 + m00 - m11 - m22, 1 - m00 + m11 - m22, 1 - m00 - m11 + m22], dim=-1)
    q_abs = torch.where(sq > 0, torch.sqrt(sq.clamp(min=0)), torch.zeros_like(sq))
    cand = torch.stack(
        [
            torch.stack([q_abs[..., 0] ** 2, m21 - m12, m02 - m20, m10 - m01], dim=-1),
            torch.stack([m21 - m12, q_abs[..., 1] ** 2, m10 + m01, m02 + m20], dim=-1),
            torch.stack([m02 - m20, m10 + m01, q_abs[..., 2] ** 2, m12 + m21], dim=-1),
            torch.stack([m10 - m01, m20 + m02, m21 + m12, q_abs[..., 3] ** 2], dim=-1),
        ],
        dim=-2,
    )
    cand = cand / (2.0 * q_abs[..., None].clamp(min=0.1))
    best = q_abs.argmax(dim=-1)
    out = torch.gather(cand, -2, best[..., None, None].expand(batch + (1, 4))).squeeze(-2)  # (w, x, y, z)
    out = out[..., [1, 2, 3, 0]]
    return torch.where(out[..., 3:4] < 0, -out, out)


def quat_inverse(q: torch.Tensor) -> torch.Tensor:
    """geometry.py:745-772."""
    conj = torch.cat([-q[..., :3], q[..., 3:]], dim=-1)
    return conj / (q * q).sum(dim=-1, keepdim=True)


def quat_multiply(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """geometry.py:775-811 (Hamilton product, scalar last)."""
    x1, y1, z1, w1 = a.unbind(-1)
    x2, y2, z2, w2 = b.unbind(-1)
    return torch.stack(
        [
            w1 * x2 + x1 * w2 + y1 * z2 - z1 * y2,
            w1 * y2 - x1 * z2 + y1 * w2 + z1 * x2,
            w1 * z2 + x1 * y2 - y1 * x2 + z1 * w2,
            w1 * w2 - x1 * x2 - y1 * y2 - z1 * z2,
        ],
        dim=-1,
    )


def relative_pose_2_to_1(q1, t1, q2, t2):
    """geometry.py:814-852 transform_pose_using_quats_and_trans_2_to_1: pose2 expressed in the frame of pose1."""
    qi = quat_inverse(q1)
    r_inv = quat_to_rotmat(qi)
    t_inv = -torch.einsum("bij,bj->bi", r_inv, t1)
    return quat_multiply(qi, q2), torch.einsum("bij,bj->bi", r_inv, t2) + t_inv


def pointmap_from_rays_depth_pose(rays, depth, trans, quats):
    """geometry.py:855-907 convert_ray_dirs_depth_along_ray_pose_trans_quats_to_pointmap.
    rays (B,H,W,3), depth (B,H,W,1), trans (B,3), quats (B,4) -> world points (B,H,W,3)."""
    quats = quats / torch.norm(quats, dim=-1, keepdim=True)
    rot = quat_to_rotmat(quats)
    b = depth.shape[0]
    pose = torch.eye(4, device=depth.device).unsqueeze(0).repeat(b, 1, 1)
    pose[:, :3, :3] = rot
    pose[:, :3, 3] = trans
    local = depth * rays
    homo = torch.cat([local, torch.ones_like(local[..., :1])], dim=-1)
    return torch.einsum("bik,bhwk->bhwi", pose, homo)[..., :3]


# ------------------------------------------------------------------------------------------------
# normalisers used by the geometric-input encoders
# ------------------------------------------------------------------------------------------------
def normalize_depth_nonzero(depth: torch.Tensor):
    """geometry.py:1523-1555 normalize_depth_using_non_zero_pixels. depth (B,H,W,1) -> (normalised, factor (B,))."""
    valid = depth > 0
    factor = (depth * valid).sum(dim=(1, 2, 3)) / (valid.sum(dim=(1, 2, 3)) + 1e-8)
    factor = factor.clip(min=1e-8)
    return depth / factor.view(-1, 1, 1, 1), factor


def normalize_pose_translations(trans: torch.Tensor):
    """geometry.py:1558-1595. trans (B,V,3) -> (normalised, factor (B,)) using the mean norm of non-zero translations."""
    dist = trans.norm(dim=-1)
    factor = dist.sum(dim=1) / ((dist > 0).sum(dim=1) + 1e-8)
    factor = factor.clip(min=1e-8)
    return trans / factor.view(-1, 1, 1), factor


def log_of_norm(x: torch.Tensor) -> torch.Tensor:
    """geometry.py:1666-1679 apply_log_to_norm: x/|x| * log1p(|x|) along the last dim."""
    n = x.norm(dim=-1, keepdim=True)
    return x / n.clip(min=1e-8) * torch.log1p(n)


# ------------------------------------------------------------------------------------------------
# numpy edge masks used by infer() post-processing
# ------------------------------------------------------------------------------------------------
def _windows3(x: np.ndarray) -> np.ndarray:
    """All 3x3 windows of a padded (H+2, W+2, ...) array, stacked on a new leading axis of size 9."""
    h, w = x.shape[0] - 2, x.shape[1] - 2
    return np.stack([x[dy : dy + h, dx : dx + w] for dy in range(3) for dx in range(3)], axis=0)


def _nanmax_pool3(x: np.ndarray) -> np.ndarray:
    """geometry.py:1905-2028 max_pool_2d(kernel 3, stride 1, padding 1): NaN padding + nanmax, applied
    separably (columns then rows is equivalent for a max)."""
    pad = np.full((x.shape[0] + 2, x.shape[1] + 2), np.nan, dtype=x.dtype)
    pad[1:-1, 1:-1] = x
    with np.errstate(all="ignore"):
        import warnings

        with warnings.catch_warnings():
            warnings.simplefilter("ignore", RuntimeWarning)
            return np.nanmax(_windows3(pad), axis=0)


def points_to_normals(points: np.ndarray, mask: np.ndarray):
    """geometry.py:1717-1780 points_to_normals (mask given, no edge_threshold). points (H,W,3), mask (H,W) bool
    -> (normals (H,W,3), normal_mask (H,W))."""
    h, w = points.shape[:2]
    mp = np.zeros((h + 2, w + 2), dtype=bool)
    mp[1:-1, 1:-1] = mask
    pp = np.zeros((h + 2, w + 2, 3), dtype=points.dtype)
    pp[1:-1, 1:-1] = points
    c = pp[1:-1, 1:-1]
    up, left, down, right = pp[:-2, 1:-1] - c, pp[1:-1, :-2] - c, pp[2:, 1:-1] - c, pp[1:-1, 2:] - c
    with np.errstate(all="ignore"):
        n = np.stack([np.cross(up, left), np.cross(left, down), np.cross(down, right), np.cross(right, up)])
        n = n / (np.linalg.norm(n, axis=-1, keepdims=True) + 1e-12)
        mu, ml, md, mr = mp[:-2, 1:-1], mp[1:-1, :-2], mp[2:, 1:-1], mp[1:-1, 2:]
        valid = np.stack([mu & ml, ml & md, md & mr, mr & mu]) & mp[None, 1:-1, 1:-1]
        n = (n * valid[..., None]).sum(axis=0)
        n = n / (np.linalg.norm(n, axis=-1, keepdims=True) + 1e-12)
    nmask = valid.any(axis=0)
    return np.where(nmask[..., None], n, 0), nmask


def depth_edge(depth: np.ndarray, rtol: float, mask: np.ndarray) -> np.ndarray:
    """geometry.py:2031-2072 depth_edge(rtol=..., mask=...): (3x3 max - 3x3 min over valid pixels) / depth > rtol."""
    with np.errstate(all="ignore"):
        diff = _nanmax_pool3(np.where(mask, depth, -np.inf)) + _nanmax_pool3(np.where(mask, -depth, -np.inf))
        return diff / depth > rtol


def normals_edge(normals: np.ndarray, tol_deg: float, mask: np.ndarray) -> np.ndarray:
    """geometry.py:2129-2188 normals_edge(tol, mask): max angle to the (edge-padded) 3x3 neighbours where the
    neighbour is valid, then a 3x3 nanmax pool, thresholded in degrees."""
    with np.errstate(all="ignore"):
        n = normals / (np.linalg.norm(normals, axis=-1, keepdims=True) + 1e-12)
        npad = np.pad(n, ((1, 1), (1, 1), (0, 0)), mode="edge")
        mpad = np.pad(mask, ((1, 1), (1, 1)), mode="edge")
        dots = (_windows3(npad) * n[None]).sum(axis=-1)  # (9,H,W), window offset (dy,dx) at index 3*dy+dx
        # Reference quirk (geometry.py:2167-2176): the mask is 2-D, so `axis=(-3,-2)` wraps to axes (1,0) and the
        # mask window comes out TRANSPOSED relative to the normals window: the angle to neighbour (dy,dx) is gated
        # by the validity of neighbour (dx,dy).  Reproduced here because the oracle must match the reference bit
        # for bit, not fix it.
        mwin = _windows3(mpad).reshape(3, 3, *mask.shape).transpose(1, 0, 2, 3).reshape(9, *mask.shape)
        ang = np.where(mwin, np.arccos(dots), 0).max(axis=0)
        ang = _nanmax_pool3(ang)
        return ang > np.deg2rad(tol_deg)


def depthmap_to_camera_frame(depthmap: torch.Tensor, intrinsics: torch.Tensor):
    """reference geometry.py:18-73: (B,H,W) depth + (B,3,3) intrinsics -> (B,H,W,3) camera-frame points, valid mask."""
    B, H, W = depthmap.shape
    xg, yg = torch.meshgrid(torch.arange(W).float(), torch.arange(H).float(), indexing="xy")
    fx, fy = intrinsics[:, 0, 0].view(-1, 1, 1), intrinsics[:, 1, 1].view(-1, 1, 1)
    cx, cy = intrinsics[:, 0, 2].view(-1, 1, 1), intrinsics[:, 1, 2].view(-1, 1, 1)
    xx = (xg[None] - cx) * depthmap / fx
    yy = (yg[None] - cy) * depthmap / fy
    return torch.stack((xx, yy, depthmap), dim=-1), depthmap > 0.0


def depthmap_to_world_frame(depthmap: torch.Tensor, intrinsics: torch.Tensor, camera_pose=None):
    """reference geometry.py:76-114."""
    pts, valid = depthmap_to_camera_frame(depthmap, intrinsics)
    if camera_pose is not None:
        homo = torch.cat([pts, torch.ones_like(pts[..., :1])], dim=-1)
        pts = torch.einsum("bik,bhwk->bhwi", camera_pose, homo)[..., :3]
    return pts, valid
