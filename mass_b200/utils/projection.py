"""Drop-in for /root/reference/mass/utils/projection.py: the five free functions
of the projection math, same names, argument order and return values, with the
tensor work done by the sm_100a kernels of libmassb200 (no CPU path).

Only the 12-float camera pose (rotation + position) is prepared on the host with
ATen CPU ops: that is what makes it bit-identical to the reference CPU path
(Sleef cos/sin, and torch.cross's fused form -- see tests/golden/pose.npz).
"""
import numpy as np
import torch

from mass_b200 import _lib

_ws = _lib.Workspace()


def spherical_to_cartesian(yaw, elevation):
    """Unit vector for a yaw (CCW from +x) and an elevation (up positive).
    Reference: mass/utils/projection.py:6-31.  Works on any device; the layers
    call it on CPU tensors so the values match the reference CPU path."""
    ce = torch.cos(elevation)
    return torch.stack([torch.cos(yaw) * ce, torch.sin(yaw) * ce, torch.sin(elevation)], dim=-1)


def project_camera_rays(image_height, image_width, focal_length_y, focal_length_x,
                        dtype=torch.float32, device='cpu'):
    """Camera-frame pinhole rays [H, W, 3] = (x, -y, -1).
    Reference: mass/utils/projection.py:34-74 (runs once per layer, on the host)."""
    rows = torch.arange(image_height, dtype=dtype, device=device)
    cols = torch.arange(image_width, dtype=dtype, device=device)
    v = (rows - 0.5 * float(image_height - 1)) / focal_length_y
    u = (cols - 0.5 * float(image_width - 1)) / focal_length_x
    vv, uu = torch.meshgrid(v, u, indexing='ij')
    return torch.stack([uu, -vv, -torch.ones_like(uu)], dim=-1)


def camera_pose(position, yaw, elevation):
    """[..., 12] float32 CPU tensor: rotation (row-major) then position, computed
    with the reference's ATen CPU op sequence (base_projection_layer.py:328-331,
    projection.py:104-105).  Accepts scalars or [T] batches."""
    yaw = torch.as_tensor(yaw, dtype=torch.float32).cpu()
    elevation = torch.as_tensor(elevation, dtype=torch.float32).cpu()
    position = torch.as_tensor(position, dtype=torch.float32).cpu()
    eye = spherical_to_cartesian(yaw, elevation)
    up = spherical_to_cartesian(yaw, elevation + np.pi / 2)
    rot = torch.stack([torch.linalg.cross(eye, up, dim=-1), up, -eye], dim=-1)
    return torch.cat([rot.reshape(*rot.shape[:-2], 9), position], dim=-1).contiguous()


def transform_rays(rays, eye_vector, up_vector):
    """Rotate camera-frame rays into the world frame.
    Reference: mass/utils/projection.py:77-110.  rays: CUDA [..., 3];
    eye/up: [3] (any device) or [B, 3] with rays [B, ..., 3]."""
    device = _lib.require_cuda(rays.device)
    rays = rays.to(torch.float32).contiguous()
    eye = torch.as_tensor(eye_vector, dtype=torch.float32).cpu()
    up = torch.as_tensor(up_vector, dtype=torch.float32).cpu()
    rot = torch.stack([torch.linalg.cross(eye, up, dim=-1), up, -eye], dim=-1)
    pose = torch.cat([rot.reshape(*rot.shape[:-2], 9), torch.zeros(*rot.shape[:-2], 3)], dim=-1)
    pose = pose.contiguous().to(device)
    out = torch.empty_like(rays)
    L = _lib.lib()
    if pose.dim() == 1:
        _lib.check(L.mb_transform_rays(_lib.stream_ptr(device), _lib.ptr(rays), rays.numel() // 3,
                                       _lib.ptr(pose), _lib.ptr(out)))
    else:
        if pose.shape[0] != rays.shape[0]:
            raise ValueError("batched eye/up need rays with the same leading batch size")
        for b in range(pose.shape[0]):
            _lib.check(L.mb_transform_rays(_lib.stream_ptr(device), _lib.ptr(rays[b]),
                                           rays[b].numel() // 3, _lib.ptr(pose[b]), _lib.ptr(out[b])))
    return out


def bin_rays(bins0, bins1, bins2, origin, rays, depth, *features,
             min_ray_depth=0.0, max_ray_depth=10.0):
    """World points -> voxel indices, in-voxel ratios and the valid subset of
    `features`, valid pixels in row-major order.
    Reference: mass/utils/projection.py:113-230 (unbatched form, as update() calls it)."""
    device = _lib.require_cuda(rays.device)
    f32 = dict(dtype=torch.float32, device=device)
    rays = rays.to(**f32).contiguous()
    depth = torch.as_tensor(depth, **f32).contiguous()
    origin = torch.as_tensor(origin, **f32).contiguous()
    bins0, bins1, bins2 = (torch.as_tensor(b, **f32).contiguous() for b in (bins0, bins1, bins2))
    npix = depth.numel()
    if rays.numel() != 3 * npix or origin.numel() != 3:
        raise ValueError("bin_rays: rays must be [..., 3] matching depth [..., 1]; origin [3]")
    L = _lib.lib()
    ind = torch.empty(4, max(npix, 1), dtype=torch.int64, device=device)
    rat = torch.empty(3, max(npix, 1), **f32)
    count = torch.zeros(1, dtype=torch.int64, device=device)
    ws = _ws.get(L.mb_bin_rays_workspace_bytes(npix), device)
    _lib.check(L.mb_bin_rays(_lib.stream_ptr(device), _lib.ptr(bins0), bins0.numel(), _lib.ptr(bins1),
                             bins1.numel(), _lib.ptr(bins2), bins2.numel(), _lib.ptr(origin),
                             _lib.ptr(rays), _lib.ptr(depth), npix, float(min_ray_depth),
                             float(max_ray_depth), _lib.ptr(ind[0]), _lib.ptr(ind[1]), _lib.ptr(ind[2]),
                             _lib.ptr(rat[0]), _lib.ptr(rat[1]), _lib.ptr(rat[2]), _lib.ptr(ind[3]),
                             _lib.ptr(count), _lib.ptr(ws), ws.numel()))
    n = int(count.item())            # same host sync as the reference's nonzero()
    pix = ind[3, :n]
    selected = [f.reshape(npix, *f.shape[depth.dim() - 1:])[pix] if f.dim() >= depth.dim()
                else f.reshape(npix)[pix] for f in features]
    return (ind[0, :n], ind[1, :n], ind[2, :n], rat[0, :n], rat[1, :n], rat[2, :n], *selected)


def update_feature_map(ind0, ind1, ind2, ratio0, ratio1, ratio2, features, feature_map,
                       interpolation_weight=1.0, exact=True):
    """Trilinear splat + per-voxel weighted-average update, IN PLACE on
    `feature_map` [S0, S1, S2, F].  Reference: mass/utils/projection.py:233-351.
    Deterministic (stable sort + per-voxel segmented reduce, no float atomics);
    exact=True reproduces the reference CPU result bit for bit."""
    device = _lib.require_cuda(feature_map.device)
    if feature_map.dtype != torch.float32 or not feature_map.is_contiguous():
        raise ValueError("feature_map must be a contiguous float32 tensor")
    if feature_map.dim() != 4:
        raise ValueError("feature_map must be [S0, S1, S2, F] (unbatched, as the layers use it)")
    S0, S1, S2, F = feature_map.shape
    i64 = dict(dtype=torch.int64, device=device)
    f32 = dict(dtype=torch.float32, device=device)
    ind0, ind1, ind2 = (torch.as_tensor(i, **i64).reshape(-1).contiguous() for i in (ind0, ind1, ind2))
    ratio0, ratio1, ratio2 = (torch.as_tensor(r, **f32).reshape(-1).contiguous()
                              for r in (ratio0, ratio1, ratio2))
    npts = ind0.numel()
    features = torch.as_tensor(features, **f32).reshape(npts, F).contiguous()
    L = _lib.lib()
    ws = _ws.get(L.mb_update_feature_map_workspace_bytes(npts, S0, S1, S2), device)
    _lib.check(L.mb_update_feature_map(
        _lib.stream_ptr(device), _lib.ptr(ind0), _lib.ptr(ind1), _lib.ptr(ind2), _lib.ptr(ratio0),
        _lib.ptr(ratio1), _lib.ptr(ratio2), _lib.ptr(features), npts, F, _lib.ptr(feature_map), S0, S1, S2,
        float(interpolation_weight), _lib.MODE_EXACT if exact else _lib.MODE_FAST, _lib.ptr(ws),
        ws.numel()))
