"""BASELINE config 3 at full size on one GPU: a furnished box-room (~200 object boxes), walkthrough (250 frames) and
unshuffle (500 frames) passes build an occupancy map, two semantic maps (class ids) and two 256-d instance-feature maps
(384x384x96 at 0.05 m, as agent.py:825-832), then the agent's matching loop (agent.py:424-450) runs
predict_scene_differences until no class differs.  Prints the time of every stage.  Inputs are synthetic and are
rendered on the GPU (input generation, not part of the measured path).   Usage: python tools/c3_episode.py"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mass_b200.nn.applications.occupancy_projection_layer import OccupancyProjectionLayer   # noqa: E402
from mass_b200.nn.applications.resnet_projection_layer import ResNetProjectionLayer         # noqa: E402
from mass_b200.nn.applications.semantic_projection_layer import SemanticProjectionLayer     # noqa: E402
from mass_b200.utils import synthetic                                                        # noqa: E402
from mass_b200.utils.experimentation import predict_scene_differences                        # noqa: E402

dev = torch.device("cuda:0")
H = W = 224
N_OBJ = 200
KW = dict(camera_height=H, camera_width=W, vertical_fov=90.0, map_height=384, map_width=384, map_depth=96,
          grid_resolution=0.05, interpolation_weight=0.5, exact=False, **synthetic.MAP_ORIGIN)


def make_boxes(shifted):
    rng = np.random.default_rng(23)
    c = rng.uniform([-3.7, -2.7, 0.0], [3.7, 2.7, 1.6], (N_OBJ, 3))
    s = rng.uniform(0.12, 0.3, (N_OBJ, 3))
    cls = rng.integers(1, 54, N_OBJ)
    if shifted:
        moved = rng.choice(N_OBJ, 25, replace=False)
        c[moved, :2] += rng.uniform(0.4, 0.8, (25, 2)) * rng.choice([-1, 1], (25, 2))
    lo, hi = c - s / 2, c + s / 2
    lo[:, 2] = np.maximum(lo[:, 2], 0.0)
    return torch.tensor(np.concatenate([lo, hi], 1), dtype=torch.float64, device=dev), torch.tensor(cls, device=dev)


def render(T, boxes, classes, feat_table):
    """depth [T,H,W,1] f32, ids [T,H,W,1] i64, features [T,56,56,256] f32 (GPU slab-method ray casting)."""
    rays = torch.tensor(synthetic.camera_rays(H, W), device=dev)                    # [H,W,3] f64
    lo_room = torch.tensor(synthetic.ROOM_LO, device=dev)
    hi_room = torch.tensor(synthetic.ROOM_HI, device=dev)
    depth, ids, feats, pos, yaw, elev = [], [], [], [], [], []
    for t in range(T):
        p, y, e = synthetic.boxroom_pose(t, T)
        rot = torch.tensor(synthetic._rotation(y, e), device=dev)
        r = rays @ rot.T
        o = torch.tensor(np.asarray(p, np.float64), device=dev)
        far = torch.where(r > 0, hi_room, lo_room)
        t_wall = ((far - o) / r).amin(-1)
        inv = 1.0 / r[..., None, :]
        t0, t1 = (boxes[:, :3] - o) * inv, (boxes[:, 3:] - o) * inv
        tn, tf = torch.minimum(t0, t1).amax(-1), torch.maximum(t0, t1).amin(-1)
        tn = torch.where((tn <= tf) & (tn > 0), tn, torch.full_like(tn, float("inf")))
        tb, k = tn.min(-1)
        hit = torch.where(tb < t_wall, k, torch.full_like(k, -1))
        d = torch.minimum(t_wall, tb).to(torch.float32)
        depth.append(d[..., None])
        ids.append(torch.where(hit >= 0, classes[hit.clamp(min=0)], torch.zeros_like(hit))[..., None])
        feats.append(feat_table[hit[2::4, 2::4] + 1])
        pos.append(p), yaw.append(y), elev.append(e)
    return dict(position=np.stack(pos), yaw=np.array(yaw, np.float32), elevation=np.array(elev, np.float32),
                depth=torch.stack(depth), semantic=torch.stack(ids), features=torch.stack(feats).contiguous())


def timed(fn):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    out = fn()
    torch.cuda.synchronize()
    return out, (time.perf_counter() - t0) * 1e3


def main():
    feat_table = torch.rand(N_OBJ + 1, 256, device=dev)
    scenes = []
    for shifted, T in ((False, 250), (True, 500)):
        boxes, classes = make_boxes(shifted)
        scenes.append(render(T, boxes, classes, feat_table))
    occ = OccupancyProjectionLayer(feature_size=1, **KW).to(dev)
    sems = [SemanticProjectionLayer(feature_size=54, **KW).to(dev) for _ in range(2)]
    ress = [ResNetProjectionLayer(feature_size=256, **KW).to(dev) for _ in range(2)]
    # warm-up at full size (the shared scratch buffer grows to its final size), then reset
    for L in (occ, sems[0], ress[0]):
        L.update_batch(scenes[1])
        L.reset(**{k: KW[k] for k in ("origin_y", "origin_x", "origin_z")})
    total = 0.0
    for i, (name, obs) in enumerate(zip(("walkthrough", "unshuffle"), scenes)):
        T = len(obs["yaw"])
        for label, layer in (("occupancy F=1", occ), ("semantic ids F=54", sems[i]), ("features 56x56 F=256", ress[i])):
            if label.startswith("occupancy") and i == 1:
                layer.reset(**{k: KW[k] for k in ("origin_y", "origin_x", "origin_z")})
            _, ms = timed(lambda: layer.update_batch(obs))
            total += ms
            print("%-11s %-22s %4d frames  %8.2f ms  (%.0f frames/s incl. host marshalling)" % (name, label, T, ms, T / ms * 1e3))
    # the agent's matching loop
    kw = dict(confidence_threshold=0.0, contour_padding=0, contour_threshold=0.0, distance_threshold=0.05)
    moved, calls, n_inst = set(), 0, 0
    t_match = 0.0
    while True:
        (obj, g0, g1), ms = timed(lambda: predict_scene_differences(sems[0], sems[1], ress[0], ress[1], moved,
                                                                     list(range(54)), **kw))
        t_match += ms
        calls += 1
        if obj is None:
            break
        n_inst += len(g0)
        moved.add(obj)
    print("matching: %d predict_scene_differences calls, %d classes differ, %d instance pairs to move, %.1f ms total "
          "(%.1f ms per call; the reference's find() alone is 0.9-1.2 s per class and map on 8 CPU threads, SURVEY.md 6)"
          % (calls, len(moved), n_inst, t_match, t_match / calls))
    total += t_match
    print("episode total (5 maps + matching): %.1f ms" % total)


if __name__ == "__main__":
    main()
