"""Generates tests/golden/*.npz by running the UNMODIFIED reference
(/root/reference, imported through tests/refshim.py) on CPU.

Run once in the build container:  python tests/golden/make_golden.py
The fixtures are committed; nothing at test time on the GPU box reads the
reference.  Large outputs are stored as SHA-256 digests of the exact bytes plus
a strided sample (indices/ratios must be bit-exact, so a digest is a complete
check; map values carry a sample for the 1e-5 tolerance check as well).
"""
import hashlib
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import refshim  # noqa: E402
from mass_b200.utils import synthetic  # noqa: E402
from make_golden_maps import build_block_maps  # noqa: E402

R = refshim.load()
P = R.projection


def digest(a):
    return np.frombuffer(hashlib.sha256(np.ascontiguousarray(a).tobytes()).digest(), np.uint8)


def save(name, **arrays):
    path = os.path.join(HERE, name)
    np.savez_compressed(path, **arrays)
    print("%-22s %8.1f KB" % (name, os.path.getsize(path) / 1024))


def ref_bin(layer, obs):
    """Runs the reference's transform_rays + bin_rays exactly as update() does."""
    position = torch.as_tensor(obs["position"], dtype=torch.float32)
    yaw = torch.as_tensor(obs["yaw"], dtype=torch.float32)
    elevation = torch.as_tensor(obs["elevation"], dtype=torch.float32)
    depth = torch.as_tensor(obs["depth"], dtype=torch.float32)
    eye = P.spherical_to_cartesian(yaw, elevation)
    up = P.spherical_to_cartesian(yaw, elevation + np.pi / 2)
    oriented = P.transform_rays(layer.rays, eye, up)
    out = P.bin_rays(layer.bins_x, layer.bins_y, layer.bins_z, position, oriented, depth)
    return [o.numpy() for o in out[:6]]


# ---------------------------------------------------------------------------
def g_kat_tiny():
    layer = R.base.BaseProjectionLayer(camera_height=2, camera_width=2, vertical_fov=90.0,
                                       map_height=8, map_width=8, map_depth=4, feature_size=2,
                                       grid_resolution=0.5, interpolation_weight=0.5)
    obs = dict(position=np.array([0.1, 0.2, 0.3], np.float32), yaw=np.float32(0.0),
               elevation=np.float32(0.0),
               depth=np.array([[1.0, 1.2], [0.0, 11.0]], np.float32)[..., None],
               features=np.array([[[1, 0], [0, 1]], [[1, 2], [3, 4]]], np.float32))
    ix, iy, iz, rx, ry, rz = ref_bin(layer, obs)
    arrays = dict(rays=layer.rays.numpy(), bins_x=layer.bins_x.numpy(), bins_y=layer.bins_y.numpy(),
                  bins_z=layer.bins_z.numpy(), ind_x=ix, ind_y=iy, ind_z=iz, ratio_x=rx, ratio_y=ry,
                  ratio_z=rz, **{"obs_" + k: np.asarray(v) for k, v in obs.items()})
    layer.update(obs)
    arrays["data1"] = layer.data.numpy().copy()
    layer.update(obs)
    arrays["data2"] = layer.data.numpy().copy()
    arrays["world_to_map"] = layer.world_to_map(torch.tensor([0.1, 0.2, 0.3])).numpy()
    arrays["map_to_world"] = layer.map_to_world(torch.tensor([4, 3, 2])).numpy()
    save("kat_tiny.npz", **arrays)


def g_pose():
    rng = np.random.default_rng(11)
    yaw = np.concatenate([[0.0, np.pi / 2, -np.pi / 2, np.pi], rng.uniform(-7, 7, 124)]).astype(np.float32)
    elev = np.concatenate([[0.0, 0.0, -np.pi / 6, 0.5], rng.uniform(-1.5, 1.5, 124)]).astype(np.float32)
    eye, up, rot = [], [], []
    for y, e in zip(yaw, elev):
        ty, te = torch.tensor(y), torch.tensor(e)
        ev = P.spherical_to_cartesian(ty, te)
        uv = P.spherical_to_cartesian(ty, te + np.pi / 2)
        eye.append(ev.numpy())
        up.append(uv.numpy())
        rot.append(torch.stack([torch.cross(ev, uv), uv, -ev], dim=-1).numpy())
    rays = torch.randn(5, 7, 3, generator=torch.Generator().manual_seed(3))
    oriented = P.transform_rays(rays, torch.as_tensor(eye[9]), torch.as_tensor(up[9])).numpy()
    save("pose.npz", yaw=yaw, elevation=elev, eye=np.stack(eye), up=np.stack(up), rot=np.stack(rot),
         rays=rays.numpy(), oriented9=oriented)


def g_small_seq():
    kw = dict(camera_height=24, camera_width=32, vertical_fov=90.0, map_height=16, map_width=20,
              map_depth=8, feature_size=5, grid_resolution=0.25, interpolation_weight=0.5,
              origin_x=0.1, origin_y=-0.2, origin_z=0.3)
    layer = R.base.BaseProjectionLayer(**kw)
    rng = np.random.default_rng(5)
    T = 6
    arrays = dict(position=rng.uniform(-1, 1, (T, 3)).astype(np.float32),
                  yaw=rng.uniform(-4, 4, T).astype(np.float32),
                  elevation=rng.uniform(-1, 1, T).astype(np.float32),
                  depth=rng.uniform(0, 4, (T, 24, 32, 1)).astype(np.float32),
                  features=rng.standard_normal((T, 24, 32, 5)).astype(np.float32))
    for t in range(T):
        obs = {k: arrays[k][t] for k in ("position", "yaw", "elevation", "depth", "features")}
        for name, a in zip(("ind_x", "ind_y", "ind_z", "ratio_x", "ratio_y", "ratio_z"), ref_bin(layer, obs)):
            arrays["%s_%d" % (name, t)] = a
        layer.update(obs)
        arrays["data_%d" % t] = layer.data.numpy().copy()
    arrays["kwargs"] = np.array(repr(kw))
    save("small_seq.npz", **arrays)


def g_lowres():
    kw = dict(camera_height=32, camera_width=32, vertical_fov=60.0, map_height=24, map_width=24,
              map_depth=10, feature_size=7, grid_resolution=0.2, interpolation_weight=0.3)
    layer = R.base.BaseProjectionLayer(**kw)
    rng = np.random.default_rng(6)
    arrays = dict(position=rng.uniform(-0.5, 0.5, (3, 3)).astype(np.float32),
                  yaw=rng.uniform(-3, 3, 3).astype(np.float32),
                  elevation=rng.uniform(-0.7, 0.7, 3).astype(np.float32),
                  depth=rng.uniform(0.2, 3, (3, 32, 32, 1)).astype(np.float32),
                  features=rng.random((3, 8, 8, 7)).astype(np.float32))
    for t in range(3):
        layer.update({k: arrays[k][t] for k in ("position", "yaw", "elevation", "depth", "features")})
    arrays["data"] = layer.data.numpy().copy()
    arrays["kwargs"] = np.array(repr(kw))
    save("lowres.npz", **arrays)


def g_border():
    """Points on/over the map border (clamped neighbours alias onto one voxel),
    special depths (0, 10, >10, <0, nan, inf) and exact bin-edge hits."""
    kw = dict(camera_height=16, camera_width=16, vertical_fov=120.0, map_height=6, map_width=6,
              map_depth=4, feature_size=3, grid_resolution=0.5, interpolation_weight=0.5)
    layer = R.base.BaseProjectionLayer(**kw)
    rng = np.random.default_rng(7)
    depth = rng.uniform(0, 3, (2, 16, 16, 1)).astype(np.float32)
    special = np.array([0.0, 10.0, 10.000001, -0.0, -1e-3, np.nan, np.inf, -np.inf, 1e-30, 9.999999], np.float32)
    depth[0, 0, :10, 0] = special
    depth[1, 5, :10, 0] = special
    arrays = dict(position=np.array([[1.2, -1.3, 0.7], [-1.49, 1.49, -0.99]], np.float32),
                  yaw=np.array([0.3, 2.0], np.float32), elevation=np.array([0.2, -0.4], np.float32),
                  depth=depth, features=rng.random((2, 16, 16, 3)).astype(np.float32))
    for t in range(2):
        obs = {k: arrays[k][t] for k in ("position", "yaw", "elevation", "depth", "features")}
        for name, a in zip(("ind_x", "ind_y", "ind_z", "ratio_x", "ratio_y", "ratio_z"), ref_bin(layer, obs)):
            arrays["%s_%d" % (name, t)] = a
        layer.update(obs)
        arrays["data_%d" % t] = layer.data.numpy().copy()
    arrays["kwargs"] = np.array(repr(kw))
    save("border.npz", **arrays)


def g_c1():
    """BASELINE config 1: one (then a second) 224x224 box-room frame, 54-class
    probabilities, 384x384x96 map at 0.05 m.  Inputs are stored (low-res
    probabilities, depth, pose); outputs as digests + strided samples."""
    kw = dict(camera_height=224, camera_width=224, vertical_fov=90.0, map_height=384, map_width=384,
              map_depth=96, feature_size=54, grid_resolution=0.05, interpolation_weight=0.5,
              **synthetic.MAP_ORIGIN)
    layer = R.base.BaseProjectionLayer(**kw)
    rays = synthetic.camera_rays(224, 224)
    arrays = dict(kwargs=np.array(repr(kw)), frames=np.array([37, 38]), num_frames=np.array(500))
    for n, t in enumerate((37, 38)):
        position, yaw, elevation = synthetic.boxroom_pose(t, 500)
        depth, _ = synthetic.render_depth(rays, position, yaw, elevation)
        low = synthetic.boxroom_probs(t, 224, 224, 54)
        obs = dict(position=position, yaw=yaw, elevation=elevation, depth=depth[..., None],
                   features=synthetic.upsample(low, 8))
        arrays.update({"position_%d" % n: position, "yaw_%d" % n: yaw, "elevation_%d" % n: elevation,
                       "depth_%d" % n: depth, "probs_low_%d" % n: low})
        names = ("ind_x", "ind_y", "ind_z", "ratio_x", "ratio_y", "ratio_z")
        for name, a in zip(names, ref_bin(layer, obs)):
            arrays["%s_sha_%d" % (name, n)] = digest(a.astype(np.int32) if a.dtype == np.int64 else a)
            arrays["%s_s_%d" % (name, n)] = a[::97].copy()
            arrays["n_valid_%d" % n] = np.array(a.shape[0])
        layer.update(obs)
        data = layer.data.numpy().reshape(-1, 54)
        occ = np.flatnonzero((data != 0).any(-1))
        arrays["occ_count_%d" % n] = np.array(occ.size)
        arrays["occ_sha_%d" % n] = digest(occ.astype(np.int64))
        sample = occ[::23]
        arrays["sample_idx_%d" % n] = sample
        arrays["sample_rows_%d" % n] = data[sample].copy()
        arrays["rows_sha_%d" % n] = digest(data[occ])
    save("c1_frames.npz", **arrays)


def g_find_match():
    S0, S1, S2, F_feat = 48, 40, 12, 16
    kw = dict(camera_height=8, camera_width=8, map_height=S0, map_width=S1, map_depth=S2,
              grid_resolution=0.05, origin_x=0.3, origin_y=-0.1, origin_z=0.9)
    layers = []
    for shift in (0, 1):
        sem, feat = build_block_maps(21, S0, S1, S2, 54, F_feat, shift)
        s = R.semantic.SemanticProjectionLayer(feature_size=54, class_to_colors=np.zeros((54, 3)), **kw)
        f = R.base.BaseProjectionLayer(feature_size=F_feat, **kw)
        s.data.copy_(torch.from_numpy(sem))
        f.data.copy_(torch.from_numpy(feat))
        layers.append((s, f))
    arrays = dict(dims=np.array([S0, S1, S2, F_feat]), kwargs=np.array(repr(kw)))
    for pad in (0, 1):
        for cls in (3, 7, 12, 20, 45, 50, 9):
            for m, (s, f) in enumerate(layers):
                conf, coord, size, feats = s.find(cls, confidence_threshold=0.0, contour_padding=pad,
                                                  contour_threshold=0.0, feature_map=f)
                tag = "p%d_c%d_m%d" % (pad, cls, m)
                n = len(conf)
                arrays["n_" + tag] = np.array(n)
                arrays["boxes_" + tag] = np.array(s.boxes, np.int64).reshape(n, 4)
                arrays["conf_" + tag] = torch.stack(conf).numpy() if n else np.zeros(0, np.float32)
                arrays["coord_" + tag] = torch.stack(coord).numpy() if n else np.zeros((0, 3), np.float32)
                arrays["size_" + tag] = torch.stack(size).numpy() if n else np.zeros(0, np.float32)
                arrays["feat_" + tag] = torch.stack(feats).numpy() if n else np.zeros((0, F_feat), np.float32)
    # predict_scene_differences: with and without feature maps, several moved sets
    cases = [(set(), list(range(54)), True), ({7}, list(range(54)), True), ({7, 12, 20}, list(range(54)), True),
             (set(), [45, 50, 3], True), (set(), list(range(54)), False), ({3, 7}, [50, 45, 20, 12], False)]
    for k, (moved, cands, use_feat) in enumerate(cases):
        (s0, f0), (s1, f1) = layers
        obj, g0, g1 = R.experimentation.predict_scene_differences(
            s0, s1, f0 if use_feat else None, f1 if use_feat else None, moved, cands,
            confidence_threshold=0.0, contour_padding=0, contour_threshold=0.0,
            distance_threshold=0.05, deformation_threshold=0.0)
        arrays["psd%d_moved" % k] = np.array(sorted(moved), np.int64)
        arrays["psd%d_cands" % k] = np.array(cands, np.int64)
        arrays["psd%d_use_feat" % k] = np.array(use_feat)
        arrays["psd%d_obj" % k] = np.array(-1 if obj is None else obj)
        arrays["psd%d_g0" % k] = torch.stack(g0).numpy() if g0 else np.zeros((0, 3), np.float32)
        arrays["psd%d_g1" % k] = torch.stack(g1).numpy() if g1 else np.zeros((0, 3), np.float32)
        if g0:  # agent.py:455-465 ordering of the returned pairs
            d = torch.norm(torch.stack(g0).unsqueeze(1) - torch.stack(g1).unsqueeze(0), dim=2)
            arrays["psd%d_order" % k] = d.amin(dim=1).argsort(descending=True).numpy()
    arrays["num_psd"] = np.array(len(cases))
    save("find_match.npz", **arrays)


def g_lsap():
    from scipy.optimize import linear_sum_assignment
    rng = np.random.default_rng(31)
    costs, rows, cols, shapes = [], [], [], []
    fixed = [np.zeros((4, 4)), np.array([[1, 1, 2], [1, 1, 2]], float), np.array([[1, 1, 2], [1, 1, 2]], float).T,
             np.ones((1, 5)), np.ones((5, 1)), np.eye(6), 1 - np.eye(5)]
    for k in range(400):
        if k < len(fixed):
            c = fixed[k]
        else:
            nr, nc = rng.integers(1, 9, 2)
            kind = k % 4
            if kind == 0:
                c = rng.integers(0, 3, (nr, nc)).astype(float)
            elif kind == 1:
                c = np.round(rng.random((nr, nc)), 1)
            elif kind == 2:
                c = rng.random((nr, nc)).astype(np.float32).astype(float)
            else:
                c = rng.integers(0, 2, (nr, nc)).astype(float)
        r, cc = linear_sum_assignment(c)
        pad = np.full((8, 8), np.nan)
        pad[:c.shape[0], :c.shape[1]] = c
        costs.append(pad)
        shapes.append(c.shape)
        rr, ccp = np.full(8, -1), np.full(8, -1)
        rr[:len(r)], ccp[:len(cc)] = r, cc
        rows.append(rr)
        cols.append(ccp)
    big = rng.random((200, 200)).astype(np.float32)
    rb, cb = linear_sum_assignment(big)
    rect = rng.random((37, 90)).astype(np.float32)
    rr2, cr2 = linear_sum_assignment(rect)
    rr3, cr3 = linear_sum_assignment(rect.T)
    save("lsap.npz", costs=np.stack(costs), shapes=np.array(shapes), rows=np.stack(rows), cols=np.stack(cols),
         big=big, big_rows=rb, big_cols=cb, rect=rect, rect_rows=rr2, rect_cols=cr2, rect_t_rows=rr3,
         rect_t_cols=cr3)


def g_pairwise():
    g = torch.Generator().manual_seed(41)
    a, b = torch.randn(23, 256, generator=g), torch.randn(31, 256, generator=g)
    d = torch.linalg.norm(a.unsqueeze(1) - b.unsqueeze(0), dim=2)
    a3, b3 = torch.randn(9, 3, generator=g), torch.randn(4, 3, generator=g)
    d3 = torch.linalg.norm(a3.unsqueeze(1) - b3.unsqueeze(0), dim=2)
    save("pairwise.npz", a=a.numpy(), b=b.numpy(), d=d.numpy(), a3=a3.numpy(), b3=b3.numpy(), d3=d3.numpy())


if __name__ == "__main__":
    torch.set_num_threads(8)
    for fn in (g_kat_tiny, g_pose, g_small_seq, g_lowres, g_border, g_c1, g_find_match, g_lsap, g_pairwise):
        fn()
