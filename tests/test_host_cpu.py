"""CPU-side checks of the product package: the C-ABI library loads and exports every
symbol include/massb200.h declares (no compute calls), the host-side pose arithmetic is
the reference's, the plain-torch helpers match the golden vectors, and the hot path
refuses to run without CUDA."""
import os
import re

import numpy as np
import pytest
import torch

from conftest import ROOT, golden


def header_symbols():
    text = open(os.path.join(ROOT, "include", "massb200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mb_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from mass_b200 import _lib
    L = _lib.lib()                       # builds with nvcc for sm_100a if missing
    names = header_symbols()
    assert len(names) >= 8
    for n in names:
        assert hasattr(L, n), "libmassb200.so does not export %s" % n
    assert set(names) == set(_lib.exported_symbols()), "ctypes table and header disagree"
    assert L.mb_version() >= 100
    assert L.mb_bin_rays_workspace_bytes(1000) > 0      # pure host arithmetic, no device needed


def test_camera_pose_matches_reference_rotation():
    from mass_b200.utils.projection import camera_pose
    g = golden("pose.npz")
    pose = camera_pose(np.zeros((len(g["yaw"]), 3), np.float32), g["yaw"], g["elevation"]).numpy()
    assert np.array_equal(pose[:, :9].reshape(-1, 3, 3), g["rot"])
    one = camera_pose(np.array([1, 2, 3], np.float32), g["yaw"][9], g["elevation"][9]).numpy()
    assert np.array_equal(one[:9].reshape(3, 3), g["rot"][9]) and one[9:].tolist() == [1, 2, 3]


def test_layer_state_and_helpers_on_cpu():
    from mass_b200.nn.base_projection_layer import BaseProjectionLayer
    from mass_b200.nn.projection_layer import ProjectionLayer
    g = golden("kat_tiny.npz")
    L = BaseProjectionLayer(camera_height=2, camera_width=2, vertical_fov=90.0, map_height=8, map_width=8,
                            map_depth=4, feature_size=2, grid_resolution=0.5, interpolation_weight=0.5)
    assert isinstance(L, ProjectionLayer) and isinstance(L, torch.nn.Module)
    assert np.array_equal(L.rays.numpy(), g["rays"])
    for b in ("bins_x", "bins_y", "bins_z"):
        assert np.array_equal(getattr(L, b).numpy(), g[b])
    assert sorted(k for k, _ in L.named_buffers()) == ["bins_x", "bins_y", "bins_z", "data", "rays"]
    assert L.world_to_map(torch.tensor([0.1, 0.2, 0.3])).tolist() == g["world_to_map"].tolist()
    assert np.array_equal(L.map_to_world(torch.tensor([4, 3, 2])).numpy(), g["map_to_world"])
    L.data.copy_(torch.from_numpy(g["data1"]))
    top = L.top_down(depth_slice=None)
    assert top.shape == (8, 8, 2) and float(top[3, 4, 0]) == float(g["data1"][3, 4, 3, 0])
    assert L.visualize(None, depth_slice=None).shape == (8, 8, 4, 3)   # as the reference: one image per z
    L.reset(origin_x=1.0, origin_y=-1.0, origin_z=0.5)
    assert float(L.data.abs().sum()) == 0.0 and abs(float(L.bins_x.mean()) - 0.75) < 1e-5


def test_no_cpu_fallback():
    from mass_b200.nn.base_projection_layer import BaseProjectionLayer
    from mass_b200.utils import projection as P
    L = BaseProjectionLayer(camera_height=4, camera_width=4, map_height=8, map_width=8, map_depth=4)
    obs = dict(position=np.zeros(3, np.float32), yaw=np.float32(0), elevation=np.float32(0),
               depth=np.ones((4, 4, 1), np.float32), features=np.ones((4, 4, 1), np.float32))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        L.update(obs)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        P.transform_rays(torch.zeros(4, 3), torch.tensor([1.0, 0, 0]), torch.tensor([0, 0, 1.0]))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        P.update_feature_map(*[torch.zeros(1, dtype=torch.int64)] * 3, *[torch.zeros(1)] * 3,
                             torch.zeros(1, 1), torch.zeros(2, 2, 2, 1))


def test_product_never_imports_oracle():
    bad = []
    for base, _, files in os.walk(os.path.join(ROOT, "mass_b200")):
        for f in files:
            if f.endswith(".py") and re.search(r"^\s*(from|import)\s+oracle\b", open(os.path.join(base, f)).read(), re.M):
                bad.append(f)
    assert not bad, bad


def test_new_entry_points_fail_loudly_without_cuda():
    """column summaries, find() and the sharded fold need the CUDA library: on a CPU layer they raise."""
    from mass_b200.nn.applications.semantic_projection_layer import SemanticProjectionLayer
    from mass_b200.nn import sharded
    from mass_b200.utils import navigation
    L = SemanticProjectionLayer(camera_height=4, camera_width=4, map_height=8, map_width=8, map_depth=4, feature_size=3)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        L.column_summary()
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        navigation.navigable_area(L)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        L.find(1, 0.0, 0, 0.0, None)
    obs = dict(position=np.zeros((1, 3), np.float32), yaw=np.zeros(1, np.float32), elevation=np.zeros(1, np.float32),
               depth=np.ones((1, 4, 4, 1), np.float32), semantic=np.ones((1, 4, 4, 1), np.int64))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        sharded.fold_frames(L, obs)


def test_layer_deepcopy_drops_transient_state():
    """A copied layer carries the buffers and settings, not the device-bound caches (graphs, events, memoised queries)."""
    import copy
    import torch
    from mass_b200.nn.applications.semantic_projection_layer import SemanticProjectionLayer
    L = SemanticProjectionLayer(camera_height=8, camera_width=8, map_height=6, map_width=6, map_depth=4, feature_size=3,
                                exact=False)
    L._find_cache = ("state", {"k": object()})
    L._edge_staging = dict(done=object())
    L.data[1, 2, 3, 0] = 2.0
    C = copy.deepcopy(L)
    assert torch.equal(C.data, L.data) and C.data.data_ptr() != L.data.data_ptr()
    assert torch.equal(C.bins_x, L.bins_x) and C.exact is False and C.feature_size == 3
    assert not hasattr(C, "_find_cache") and not hasattr(C, "_edge_staging") and C._frame_graphs == {}
    C.reset(origin_x=0.5)
    assert float(C.data.abs().sum()) == 0.0 and float(L.data.abs().sum()) == 2.0
