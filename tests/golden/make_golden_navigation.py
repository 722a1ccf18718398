"""Generates tests/golden/navigation.npz with the UNMODIFIED reference: batched world_to_map / map_to_world of
BaseProjectionLayer (mass/nn/base_projection_layer.py:452-547) and NavigationPolicy's navigable_area /
reset_navigation_graph / update_navigation_graph (mass/navigation_policy.py:173-341) driven through a stand-in `self`
(the class constructor needs the simulator task; the three methods only read `feature_maps` and `navigation_graph`).
Run once in the build container:  python tests/golden/make_golden_navigation.py"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import refshim  # noqa: E402

R = refshim.load()
import mass.navigation_policy as navpol  # noqa: E402

KW = dict(camera_height=8, camera_width=8, map_height=44, map_width=52, map_depth=12, feature_size=3,
          grid_resolution=0.1, origin_x=0.35, origin_y=-0.2, origin_z=0.5)


def main():
    rng = np.random.default_rng(77)
    layer = R.base.BaseProjectionLayer(**KW)
    world = rng.uniform(-4, 4, (300, 3)).astype(np.float32)
    world[:5] = [[0.35, -0.2, 0.5], [100, 100, 100], [-100, -100, -100], [0.4, -0.25, 0.55], [2.95, 2.0, 1.1]]
    world[5:45, 0] = layer.bins_x.numpy()[:40]                   # exactly on edges
    world[45:85, 1] = layer.bins_y.numpy()[:40]
    cells3 = layer.world_to_map(torch.from_numpy(world)).numpy()
    cells2 = layer.world_to_map(torch.from_numpy(world[:, :2])).numpy()
    mapc = rng.uniform(-3, 56, (300, 3)).astype(np.float32)
    mapc[:40] = np.floor(mapc[:40])                              # integer cells
    back3 = layer.map_to_world(torch.from_numpy(mapc)).numpy()
    # (the reference's clamp_to_map cannot take xy vectors: view(1, 3) of a 2-element bound, line 449)
    # a map with obstacles: random occupied voxels + a wall
    data = np.zeros((44, 52, 12, 3), np.float32)
    occ = rng.random((44, 52, 12)) < 0.004
    data[occ] = rng.random((int(occ.sum()), 3)).astype(np.float32)
    data[20, 10:40, 2:5] = 0.5
    layer.data = torch.from_numpy(data.copy())
    me = types.SimpleNamespace(feature_maps={"nav": layer}, navigation_graph=None)
    me.navigable_area = types.MethodType(navpol.NavigationPolicy.navigable_area, me)
    out = dict(kwargs=str(KW), world=world, cells3=cells3, cells2=cells2, mapc=mapc, back3=back3, data=data)
    for tag, (pad, sl, thr, step) in {"a": (3, None, 0.0, 5), "b": (1, slice(1, 6), 0.3, 4)}.items():
        nav = navpol.NavigationPolicy.navigable_area(me, "nav", padding=pad, depth_slice=sl, obstacle_threshold=thr)
        navpol.NavigationPolicy.reset_navigation_graph(me, "nav", step_size=step, padding=pad, depth_slice=sl,
                                                       obstacle_threshold=thr)
        edges = np.array([[a[0], a[1], b[0], b[1]] for a, b in me.navigation_graph.edges()], np.int64)
        nodes = np.array(list(me.navigation_graph.nodes()), np.int64)
        # new obstacles appear, then the refresh
        data2 = data.copy()
        data2[8:12, 30:33, 2:4] = 0.7
        data2[30, 5:25, 3] = 0.9
        layer.data = torch.from_numpy(data2)
        navpol.NavigationPolicy.update_navigation_graph(me, "nav", padding=pad, depth_slice=sl, obstacle_threshold=thr)
        edges2 = np.array([[a[0], a[1], b[0], b[1]] for a, b in me.navigation_graph.edges()], np.int64).reshape(-1, 4)
        nodes2 = np.array(list(me.navigation_graph.nodes()), np.int64).reshape(-1, 2)
        layer.data = torch.from_numpy(data.copy())
        out.update({"nav_" + tag: nav.numpy(), "edges_" + tag: edges, "nodes_" + tag: nodes, "edges2_" + tag: edges2,
                    "nodes2_" + tag: nodes2, "args_" + tag: np.array([pad, -1 if sl is None else sl.start,
                                                                         -1 if sl is None else sl.stop, step], np.int64),
                    "thr_" + tag: np.float32(thr)})
    out["data2"] = data2
    path = os.path.join(HERE, "navigation.npz")
    np.savez_compressed(path, **out)
    print("navigation.npz %.1f KB" % (os.path.getsize(path) / 1024), {k: v.shape for k, v in out.items() if hasattr(v, "shape") and k.startswith(("edges", "nodes"))})


if __name__ == "__main__":
    main()
