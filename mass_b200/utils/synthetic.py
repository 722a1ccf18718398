"""Deterministic synthetic RGB-D inputs that stand in for the AI2-THOR simulator,
Mask R-CNN and ResNet-50 (all out of scope).

"Box-room": an empty axis-aligned room seen by a camera that circles inside it,
optionally furnished with axis-aligned object boxes carrying a class id.  Depth
is the planar z-depth the reference expects (camera-frame ray has z = -1, so the
world point is ``position + oriented_ray * depth``,
/root/reference/mass/utils/projection.py:182).

Everything here is input generation: float64 numpy on the host, no CUDA, no
dependence on the kernels under test.
"""
import math

import numpy as np
import torch

ROOM_LO = np.array([-4.0, -3.0, 0.0])
ROOM_HI = np.array([4.0, 3.0, 2.6])
MAP_ORIGIN = dict(origin_x=0.0, origin_y=0.0, origin_z=0.9)


def camera_rays(height, width, vertical_fov=90.0):
    """float64 pinhole rays [H, W, 3] = (x, -y, -1) / f (same model as
    project_camera_rays; used only to render synthetic depth)."""
    f = height / 2.0 / math.tan(math.radians(vertical_fov) / 2.0)
    y, x = np.meshgrid(np.arange(height, dtype=np.float64),
                       np.arange(width, dtype=np.float64), indexing="ij")
    return np.stack([(x - 0.5 * (width - 1)) / f, -(y - 0.5 * (height - 1)) / f,
                     -np.ones_like(x)], axis=-1)


def boxroom_pose(t, num_frames):
    ang = 2.0 * math.pi * t / num_frames
    position = np.array([1.5 * math.cos(ang), 1.0 * math.sin(ang), 0.9], np.float32)
    yaw = np.float32((3.0 * ang) % (2.0 * math.pi))
    elevation = np.float32(-math.pi / 6.0)
    return position, yaw, elevation


def _rotation(yaw, elevation):
    def s2c(y, e):
        return np.array([math.cos(y) * math.cos(e), math.sin(y) * math.cos(e), math.sin(e)])
    eye, up = s2c(float(yaw), float(elevation)), s2c(float(yaw), float(elevation) + math.pi / 2)
    return np.stack([np.cross(eye, up), up, -eye], axis=-1)


def render_depth(rays, position, yaw, elevation, boxes=None):
    """Planar depth [H, W] float32 of the room walls (and optional boxes, an
    [n, 6] array of (lo_xyz, hi_xyz)); also returns the index of the box hit
    (-1 = wall) per pixel."""
    rot = _rotation(yaw, elevation)
    r = rays @ rot.T                                     # oriented rays, world frame
    pos = np.asarray(position, np.float64)
    with np.errstate(divide="ignore", invalid="ignore"):
        far = np.where(r > 0, ROOM_HI, ROOM_LO)
        t_wall = np.where(r != 0, (far - pos) / r, np.inf).min(axis=-1)
        depth, hit = t_wall, np.full(t_wall.shape, -1, np.int64)
        if boxes is not None and len(boxes):
            b = np.asarray(boxes, np.float64)
            inv = 1.0 / r[..., None, :]                                  # [H,W,1,3]
            t0 = (b[None, None, :, :3] - pos) * inv
            t1 = (b[None, None, :, 3:] - pos) * inv
            tn = np.minimum(t0, t1).max(axis=-1)
            tf = np.maximum(t0, t1).min(axis=-1)
            ok = (tn <= tf) & (tn > 0)
            tn = np.where(ok, tn, np.inf)
            k = tn.argmin(axis=-1)
            tb = np.take_along_axis(tn, k[..., None], axis=-1)[..., 0]
            hit = np.where(tb < depth, k, -1)
            depth = np.minimum(depth, tb)
    return depth.astype(np.float32), hit


def boxroom_probs(t, height, width, feature_size, down=8):
    """softmax(4 * randn[h/down, w/down, F]) at LOW resolution (seed 1000 + t);
    nearest up-sampling by ``down`` gives the per-pixel class probabilities."""
    g = torch.Generator().manual_seed(1000 + t)
    z = torch.randn(height // down, width // down, feature_size, generator=g)
    return torch.softmax(4.0 * z, dim=-1).numpy()


def upsample(features, factor):
    return np.repeat(np.repeat(features, factor, axis=0), factor, axis=1)


def boxroom_features(t, height, width, feature_size):
    """rand[h, w, F] stand-in for ResNet-50 layer1 activations (seed 2000 + t)."""
    g = torch.Generator().manual_seed(2000 + t)
    return torch.rand(height, width, feature_size, generator=g).numpy()


def boxroom_frame(t, num_frames, height=224, width=224, feature_size=54, vertical_fov=90.0,
                  rays=None, down=8):
    """One observation dict in the reference's convention
    (/root/reference/mass/nn/base_projection_layer.py:309-325)."""
    rays = camera_rays(height, width, vertical_fov) if rays is None else rays
    position, yaw, elevation = boxroom_pose(t, num_frames)
    depth, _ = render_depth(rays, position, yaw, elevation)
    probs = upsample(boxroom_probs(t, height, width, feature_size, down), down)
    return dict(position=position, yaw=yaw, elevation=elevation,
                depth=depth[..., None], features=np.ascontiguousarray(probs))
