"""oracle/oracle.py -- TEST INFRASTRUCTURE ONLY (never imported by mass_b200/).

ctypes front-end of ``oracle/mass_oracle.c`` plus numpy restatements of the
host-side parts of the mapping-and-matching path of brandontrabucco/mass.
Importers allowed: tests/, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs.

Parity pin: tests/golden/*.npz (generated from the unmodified reference by
tests/golden/make_golden.py) and, when /root/reference is present, the
differential tests in tests/test_oracle_vs_reference.py.

Reference citations are relative to /root/reference.
"""
import ctypes
import os
import subprocess

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

_f32p = ctypes.POINTER(ctypes.c_float)
_f64p = ctypes.POINTER(ctypes.c_double)
_i64p = ctypes.POINTER(ctypes.c_int64)


def build(force=False):
    """Compile the C restatement with the committed Makefile (gcc only)."""
    so = os.path.join(_HERE, "libmass_oracle.so")
    src = os.path.join(_HERE, "mass_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = ctypes.CDLL(build())
        L.orc_ws_create.restype = ctypes.c_void_p
        L.orc_ws_create.argtypes = [ctypes.c_int64]
        L.orc_ws_destroy.argtypes = [ctypes.c_void_p]
        L.orc_bin_rays.restype = ctypes.c_int64
        L.orc_update_feature_map.restype = ctypes.c_int64
        L.orc_layer_update.restype = ctypes.c_int64
        L.orc_lsap.restype = ctypes.c_int
        _LIB = L
    return _LIB


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _p(a, t=_f32p):
    return a.ctypes.data_as(t)


# -- a1: mass/utils/projection.py:6-31 ---------------------------------------
def spherical_to_cartesian(yaw, elevation):
    """ATen CPU cos/sin are the reference's own arithmetic here (Sleef, not
    libm: numpy's cosf differs in the last bit on ~15 % of inputs), so this one
    function stays on torch CPU ops.  Accepts scalars or [T] arrays."""
    yaw = torch.as_tensor(yaw, dtype=torch.float32, device="cpu")
    elevation = torch.as_tensor(elevation, dtype=torch.float32, device="cpu")
    return torch.stack([torch.cos(yaw) * torch.cos(elevation),
                        torch.sin(yaw) * torch.cos(elevation),
                        torch.sin(elevation)], dim=-1).numpy()


def eye_up(yaw, elevation):
    """eye/up pair of mass/nn/base_projection_layer.py:328-331 (elevation + pi/2
    is an fp32 tensor plus a Python double, i.e. an fp32 add of fp32(pi/2))."""
    elevation = torch.as_tensor(elevation, dtype=torch.float32, device="cpu")
    return (spherical_to_cartesian(yaw, elevation),
            spherical_to_cartesian(yaw, elevation + np.pi / 2))


# -- a2 ------------------------------------------------------------------------
def focal_length(camera_height, vertical_fov):
    # mass/nn/base_projection_layer.py:151-152
    return camera_height / 2.0 / np.tan(np.radians(vertical_fov) / 2.0)


def project_camera_rays(height, width, focal_y, focal_x):
    rays = np.empty((height, width, 3), np.float32)
    lib().orc_project_camera_rays(ctypes.c_int(height), ctypes.c_int(width),
                                  ctypes.c_double(focal_y), ctypes.c_double(focal_x), _p(rays))
    return rays


# -- a3 ------------------------------------------------------------------------
def rotation_from_eye_up(eye, up):
    eye, up = _f32(eye), _f32(up)
    rot = np.empty(9, np.float32)
    lib().orc_rotation_from_eye_up(_p(eye), _p(up), _p(rot))
    return rot.reshape(3, 3)


def transform_rays(rays, eye, up):
    rays = _f32(rays)
    rot = _f32(rotation_from_eye_up(eye, up))
    out = np.empty_like(rays)
    lib().orc_transform_rays(_p(rays), ctypes.c_int64(rays.size // 3), _p(rot), _p(out))
    return out


# -- a4 ------------------------------------------------------------------------
def bin_rays(bins0, bins1, bins2, origin, rays, depth, min_ray_depth=0.0, max_ray_depth=10.0):
    """Returns (ind0, ind1, ind2, ratio0, ratio1, ratio2, pix) for the valid
    pixels in row-major order; pix are flat pixel ids (features[pix] is the
    reference's ``features[indices]``)."""
    bins0, bins1, bins2 = _f32(bins0), _f32(bins1), _f32(bins2)
    origin, rays, depth = _f32(origin), _f32(rays), _f32(depth)
    npix = depth.size
    ind = np.empty((4, npix), np.int64)
    rat = np.empty((3, npix), np.float32)
    n = lib().orc_bin_rays(_p(bins0), ctypes.c_int(bins0.size), _p(bins1), ctypes.c_int(bins1.size),
                           _p(bins2), ctypes.c_int(bins2.size), _p(origin), _p(rays), _p(depth),
                           ctypes.c_int64(npix), ctypes.c_float(min_ray_depth),
                           ctypes.c_float(max_ray_depth),
                           _p(ind[0], _i64p), _p(ind[1], _i64p), _p(ind[2], _i64p),
                           _p(rat[0]), _p(rat[1]), _p(rat[2]), _p(ind[3], _i64p))
    return (ind[0, :n].copy(), ind[1, :n].copy(), ind[2, :n].copy(),
            rat[0, :n].copy(), rat[1, :n].copy(), rat[2, :n].copy(), ind[3, :n].copy())


class Workspace:
    def __init__(self, nvox):
        self.nvox = int(nvox)
        self.h = ctypes.c_void_p(lib().orc_ws_create(ctypes.c_int64(self.nvox)))

    def __del__(self):
        try:
            lib().orc_ws_destroy(self.h)
        except Exception:
            pass


# -- a5 ------------------------------------------------------------------------
def update_feature_map(ind0, ind1, ind2, ratio0, ratio1, ratio2, features, feature_map,
                       interpolation_weight=1.0, ws=None, nthreads=1):
    """In place on ``feature_map`` (C-contiguous float32 [S0,S1,S2,F]).
    Returns the number of touched voxels."""
    assert feature_map.dtype == np.float32 and feature_map.flags.c_contiguous
    S0, S1, S2, F = feature_map.shape
    ws = ws or Workspace(S0 * S1 * S2)
    i0, i1, i2 = (np.ascontiguousarray(a, dtype=np.int64) for a in (ind0, ind1, ind2))
    r0, r1, r2 = _f32(ratio0), _f32(ratio1), _f32(ratio2)
    feats = _f32(features).reshape(-1, F)
    return lib().orc_update_feature_map(
        ws.h, _p(i0, _i64p), _p(i1, _i64p), _p(i2, _i64p), _p(r0), _p(r1), _p(r2), _p(feats),
        ctypes.c_int64(i0.size), ctypes.c_int(F), _p(feature_map), ctypes.c_int(S0),
        ctypes.c_int(S1), ctypes.c_int(S2), ctypes.c_float(interpolation_weight),
        ctypes.c_int(nthreads))


def make_bins(origin, size, resolution):
    """mass/nn/base_projection_layer.py:164-181: torch.arange on CPU, fp32.  The
    edge table is ATen's (not reproducible as min + i*step): take it from torch."""
    hi = origin + (size + 1) * resolution / 2 - 1e-6
    lo = origin - (size + 1) * resolution / 2
    return torch.arange(lo, hi, resolution, dtype=torch.float32).numpy()


class OracleLayer:
    """State + update of BaseProjectionLayer (mass/nn/base_projection_layer.py:
    67-181, 183-235, 282-343) on numpy arrays, driven by the C restatement."""

    def __init__(self, camera_height=224, camera_width=224, vertical_fov=90.0, map_height=256,
                 map_width=256, map_depth=64, feature_size=1, origin_y=0.0, origin_x=0.0,
                 origin_z=0.0, grid_resolution=0.05, interpolation_weight=0.5, nthreads=1):
        self.camera_height, self.camera_width = camera_height, camera_width
        self.map_height, self.map_width, self.map_depth = map_height, map_width, map_depth
        self.feature_size = feature_size
        self.grid_resolution = grid_resolution
        self.interpolation_weight = interpolation_weight
        self.nthreads = nthreads
        f = focal_length(camera_height, vertical_fov)
        self.rays = project_camera_rays(camera_height, camera_width, f, f)
        self.data = np.zeros((map_height, map_width, map_depth, feature_size), np.float32)
        self.ws = Workspace(map_height * map_width * map_depth)
        self.n_valid = 0
        self.n_touched = 0
        self._set_origin(origin_y, origin_x, origin_z)

    def _set_origin(self, origin_y, origin_x, origin_z):
        self.origin_y, self.origin_x, self.origin_z = origin_y, origin_x, origin_z
        self.bins_x = make_bins(origin_x, self.map_width, self.grid_resolution)
        self.bins_y = make_bins(origin_y, self.map_height, self.grid_resolution)
        self.bins_z = make_bins(origin_z, self.map_depth, self.grid_resolution)

    def reset(self, origin_y=0.0, origin_x=0.0, origin_z=0.0):
        self.data[...] = 0
        self._set_origin(origin_y, origin_x, origin_z)

    def update(self, observation):
        position = _f32(observation["position"])
        eye, up = eye_up(observation["yaw"], observation["elevation"])
        rot = _f32(rotation_from_eye_up(eye, up))
        depth = _f32(observation["depth"]).reshape(self.camera_height, self.camera_width)
        feats = _f32(observation["features"])
        fh, fw, F = feats.shape
        assert F == self.feature_size
        nv = ctypes.c_int64(0)
        self.n_touched = lib().orc_layer_update(
            self.ws.h, _p(self.rays), _p(depth), _p(feats), ctypes.c_int(self.camera_height),
            ctypes.c_int(self.camera_width), ctypes.c_int(fh), ctypes.c_int(fw), ctypes.c_int(F),
            _p(rot), _p(position), _p(self.bins_x), ctypes.c_int(self.bins_x.size),
            _p(self.bins_y), ctypes.c_int(self.bins_y.size), _p(self.bins_z),
            ctypes.c_int(self.bins_z.size), _p(self.data),
            ctypes.c_float(self.interpolation_weight), ctypes.c_float(0.0), ctypes.c_float(10.0),
            ctypes.c_int(self.nthreads), ctypes.byref(nv))
        self.n_valid = nv.value
        return self

    # a10: mass/nn/base_projection_layer.py:452-511 restricted to integer cells
    # (the (r-l)*frac term is exactly zero there): bin mid-points, y flipped.
    def cell_centres(self):
        mx = (self.bins_x[:-1] + self.bins_x[1:]) / np.float32(2)
        my = ((self.bins_y[:-1] + self.bins_y[1:])[::-1]) / np.float32(2)
        mz = (self.bins_z[:-1] + self.bins_z[1:]) / np.float32(2)
        return mx.astype(np.float32), my.astype(np.float32), mz.astype(np.float32)


# -- a11: mass/nn/applications/semantic_projection_layer.py:257-362 -----------
def _box_mean3d(mask, pad):
    """avg_pool3d(kernel 2p+1, stride 1, zero padding p, count_include_pad):
    separable running sums in float64, one division by k^3."""
    if pad == 0:
        return mask
    k = 2 * pad + 1
    out = mask.astype(np.float64)
    for axis in range(3):
        padded = np.pad(out, [(pad, pad) if a == axis else (0, 0) for a in range(3)])
        c = np.cumsum(padded, axis=axis)
        c = np.concatenate([np.zeros_like(np.take(c, [0], axis=axis)), c], axis=axis)
        n = out.shape[axis]
        hi = np.take(c, np.arange(k, k + n), axis=axis)
        lo = np.take(c, np.arange(0, n), axis=axis)
        out = hi - lo
    return (out / float(k ** 3)).astype(np.float32)


def class_presence(data, semantic_category, contour_padding=3, contour_threshold=0.0):
    """uint8 [S0,S1] image of semantic_projection_layer.py:309-317."""
    smooth = _box_mean3d(data[..., semantic_category], contour_padding)
    return (smooth > np.float32(contour_threshold)).any(axis=2).astype(np.uint8)


def find_boxes(threshold_image):
    """semantic_projection_layer.py:323-328: OpenCV (third party, as in the
    reference) decides instance count and order."""
    import cv2
    contours = cv2.findContours(np.ascontiguousarray(threshold_image), cv2.RETR_LIST,
                                cv2.CHAIN_APPROX_SIMPLE)[0]
    return [tuple(int(v) for v in cv2.boundingRect(c)) for c in contours]


def find(layer, semantic_category, confidence_threshold=0.2, contour_padding=3,
         contour_threshold=0.0, feature_map=None):
    """Returns (confidences, coordinates, sizes, features-or-None, boxes); sums
    are accumulated in float64 and rounded once (the reference's fp32 pairwise
    sums are not restated; tests allow 1e-5 relative)."""
    data = layer.data
    mask = data[..., semantic_category].astype(np.float64)
    mx, my, mz = layer.cell_centres()
    img = class_presence(data, semantic_category, contour_padding, contour_threshold)
    confs, coords, sizes, feats, boxes = [], [], [], [], []
    for (x, y, w, h) in find_boxes(img):
        roi = mask[y:y + h, x:x + w]
        total = roi.sum()
        weights = roi / (total + 1e-9)
        conf = (roi * weights).sum()
        if not (np.float32(conf) > np.float32(confidence_threshold)):
            continue
        boxes.append((x, y, w, h))
        confs.append(np.float32(conf))
        cx = (weights.sum(axis=(0, 2)) * mx[x:x + w].astype(np.float64)).sum()
        cy = (weights.sum(axis=(1, 2)) * my[y:y + h].astype(np.float64)).sum()
        cz = (weights.sum(axis=(0, 1)) * mz.astype(np.float64)).sum()
        coords.append(np.array([cx, cy, cz], np.float32))
        sizes.append(np.float32(total))
        if feature_map is not None:
            froi = feature_map.data[y:y + h, x:x + w].astype(np.float64)
            feats.append(np.tensordot(weights, froi, axes=([0, 1, 2], [0, 1, 2])).astype(np.float32))
    return confs, coords, sizes, (feats if feature_map is not None else None), boxes


# -- a12 -----------------------------------------------------------------------
# class tables: mass/thor/segmentation_config.py:43-117 (id 0 = OccupiedSpace,
# 1..43 pickable, 44..53 openable)
NUM_CLASSES = 54
ID_TO_PICKABLE = [1 <= i <= 43 for i in range(NUM_CLASSES)]
ID_TO_OPENABLE = [44 <= i <= 53 for i in range(NUM_CLASSES)]


def pairwise_l2(a, b):
    a, b = _f32(a), _f32(b)
    out = np.empty((a.shape[0], b.shape[0]), np.float32)
    lib().orc_pairwise_l2(_p(a), ctypes.c_int(a.shape[0]), _p(b), ctypes.c_int(b.shape[0]),
                          ctypes.c_int(a.shape[1]), _p(out))
    return out


def lsap(cost):
    """scipy.optimize.linear_sum_assignment restated (rows, cols as int64)."""
    cost = np.ascontiguousarray(cost, dtype=np.float64)
    nr, nc = cost.shape
    k = min(nr, nc)
    rows, cols = np.empty(k, np.int64), np.empty(k, np.int64)
    rc = lib().orc_lsap(_p(cost, _f64p), ctypes.c_int(nr), ctypes.c_int(nc),
                        _p(rows, _i64p), _p(cols, _i64p))
    if rc != 0:
        raise ValueError("cost matrix is infeasible")
    return rows, cols


def predict_scene_differences(sem0, sem1, res0, res1, objects_moved, object_ids_to_move_pred,
                              confidence_threshold=0.2, contour_padding=3, contour_threshold=0.0,
                              distance_threshold=0.0, deformation_threshold=0.0):
    """mass/utils/experimentation.py:169-313 on OracleLayer maps.  Also returns
    the per-class assignment that was examined last (rows, cols) for parity
    checks of the match indices."""
    object_to_move, goals0, goals1, last = None, [], [], None
    for cand in object_ids_to_move_pred:
        pick, opn = ID_TO_PICKABLE[cand], ID_TO_OPENABLE[cand]
        if cand in objects_moved or not (pick or opn):
            continue
        kw = dict(confidence_threshold=confidence_threshold, contour_padding=contour_padding,
                  contour_threshold=contour_threshold)
        c0, g0, s0, f0, _ = find(sem0, cand, feature_map=res0, **kw)
        c1, g1, s1, f1, _ = find(sem1, cand, feature_map=res1, **kw)
        if len(c0) == 0 or len(c1) == 0:
            continue
        if f0 is not None and f1 is not None:
            deformation = pairwise_l2(np.stack(f0), np.stack(f1))
        else:
            deformation = np.abs(np.stack(s0)[:, None] - np.stack(s1)[None, :]).astype(np.float32)
        g0, g1 = np.stack(g0), np.stack(g1)
        distance = pairwise_l2(g0, g1)
        rows, cols = lsap(deformation if pick else distance)
        last = (cand, rows.copy(), cols.copy())
        for i, j in zip(rows, cols):
            if (pick and distance[i, j] > np.float32(distance_threshold)) or opn:
                object_to_move = cand
                goals0.append(g0[i])
                goals1.append(g1[j])
        if object_to_move is not None:
            break
    return object_to_move, goals0, goals1, last


# -- next to the path ------------------------------------
def world_to_map(layer, coords):
    """mass/nn/base_projection_layer.py:513-547 with clamp_to_world (:381-413): fp32 clamp to the span of the voxel
    mid-points, searchsorted(side='right') - 1 on the layer's own edge tables, y index flipped.  coords [..., 2|3]."""
    coords = _f32(coords)
    k = coords.shape[-1]
    out = np.empty(coords.shape, np.int64)
    for axis, bins in enumerate((layer.bins_x, layer.bins_y, layer.bins_z)[:k]):
        lo = (bins[0] + bins[1]) / np.float32(2)
        hi = (bins[-1] + bins[-2]) / np.float32(2)
        x = np.minimum(np.maximum(coords[..., axis], lo), hi)
        idx = np.searchsorted(bins, x, side="right") - 1
        out[..., axis] = (bins.size - 2 - idx) if axis == 1 else idx
    return out


def map_to_world(layer, coords):
    """mass/nn/base_projection_layer.py:452-511 with clamp_to_map (:415-450), every fp32 operation rounded on its own:
    left + (right - left) * (coords - floor(coords)) between neighbouring voxel mid-points."""
    coords = _f32(coords)
    k = coords.shape[-1]
    out = np.empty(coords.shape, np.float32)
    mids = layer.cell_centres()
    sizes = (layer.map_width, layer.map_height, layer.map_depth)
    for axis in range(k):
        x = np.minimum(np.maximum(coords[..., axis], np.float32(0)), np.float32(sizes[axis] - 1))
        fl = np.floor(x)
        i = fl.astype(np.int64)
        left = mids[axis][i]
        right = mids[axis][np.minimum(np.maximum(i + 1, 0), sizes[axis] - 1)]
        out[..., axis] = (left + ((right - left).astype(np.float32) * (x - fl).astype(np.float32)).astype(np.float32)).astype(np.float32)
    return out


def navigable_area(layer, padding=3, depth_slice=None, obstacle_threshold=0.0):
    """mass/navigation_policy.py:205-221: cells with no occupied voxel in the depth slice, obstacles grown by
    `padding` cells (max_pool2d pads with -inf: the image border does not count as an obstacle)."""
    occ = np.abs(layer.data.astype(np.float64)).sum(axis=3).astype(np.float32) > np.float32(obstacle_threshold)
    if depth_slice is not None:
        occ = occ[:, :, depth_slice]
    blocked = occ.any(axis=2)
    S0, S1 = blocked.shape
    grown = np.zeros_like(blocked)
    for dy in range(-padding, padding + 1):
        for dx in range(-padding, padding + 1):
            ys, yd = slice(max(dy, 0), S0 + min(dy, 0)), slice(max(-dy, 0), S0 + min(-dy, 0))
            xs, xd = slice(max(dx, 0), S1 + min(dx, 0)), slice(max(-dx, 0), S1 + min(-dx, 0))
            grown[yd, xd] |= blocked[ys, xs]
    return (~grown).astype(np.float32)


def navigation_graph_edges(layer, navigable, step_size=5):
    """mass/navigation_policy.py:253-285: the edge list of reset_navigation_graph, in its insertion order, as
    ((x0, y0), (x1, y1)) tuples."""
    origin = np.array([[layer.origin_x, layer.origin_y]], np.float32)
    ox, oy = (int(v) % step_size for v in world_to_map(layer, origin)[0])
    S0, S1 = navigable.shape
    edges = []
    for i in range(oy, S0, step_size):
        for j in range(ox, S1, step_size):
            for di, dj in ((step_size, 0), (0, step_size)):
                y, x = i + di, j + dj
                if 0 <= y < S0 and 0 <= x < S1 and (navigable[min(i, y):max(i, y) + 1, min(j, x):max(j, x) + 1] == 1).all():
                    edges.append(((j, i), (x, y)))
    return edges


def update_navigation_graph(nodes, edges, navigable):
    """mass/navigation_policy.py:315-341 on plain lists: the nodes and edges that survive the refresh."""
    keep_nodes = [(j, i) for (j, i) in nodes if navigable[i, j] != 0]
    alive = set(keep_nodes)
    keep_edges = [((j, i), (x, y)) for (j, i), (x, y) in edges
                  if (j, i) in alive and (x, y) in alive and
                  not (navigable[min(i, y):max(i, y) + 1, min(j, x):max(j, x) + 1] == 0).any()]
    return keep_nodes, keep_edges
