"""Differential tests: CPU oracle vs the LIVE reference.  Runs only where
/root/reference exists (the build container); skipped on the GPU box."""
import numpy as np
import pytest
import torch

import refshim

pytestmark = pytest.mark.skipif(not refshim.available(), reason="reference tree not present")


@pytest.fixture(scope="module")
def ref():
    return refshim.load()


@pytest.mark.parametrize("seed", range(4))
def test_random_layers_bitwise(oracle, ref, seed):
    rng = np.random.default_rng(100 + seed)
    H, W = int(rng.integers(4, 40)), int(rng.integers(4, 40))
    kw = dict(camera_height=H, camera_width=W, vertical_fov=float(rng.uniform(40, 120)),
              map_height=int(rng.integers(4, 30)), map_width=int(rng.integers(4, 30)),
              map_depth=int(rng.integers(2, 12)), feature_size=int(rng.integers(1, 9)),
              grid_resolution=float(rng.choice([0.05, 0.1, 0.25, 0.33])),
              interpolation_weight=float(rng.choice([0.5, 0.3, 1.0])),
              origin_x=float(rng.uniform(-1, 1)), origin_y=float(rng.uniform(-1, 1)),
              origin_z=float(rng.uniform(-1, 1)))
    a, b = ref.base.BaseProjectionLayer(**kw), oracle.OracleLayer(**kw)
    assert np.array_equal(a.rays.numpy(), b.rays)
    assert np.array_equal(a.bins_x.numpy(), b.bins_x) and np.array_equal(a.bins_z.numpy(), b.bins_z)
    ext = kw["grid_resolution"] * kw["map_width"]
    for t in range(5):
        obs = dict(position=rng.uniform(-ext / 3, ext / 3, 3).astype(np.float32),
                   yaw=np.float32(rng.uniform(-7, 7)), elevation=np.float32(rng.uniform(-1.5, 1.5)),
                   depth=rng.uniform(0, ext, (H, W, 1)).astype(np.float32),
                   features=rng.standard_normal((H, W, kw["feature_size"])).astype(np.float32))
        a.update(obs)
        b.update(obs)
        assert np.array_equal(a.data.numpy(), b.data), t
    a.reset(origin_x=0.5, origin_y=0.25, origin_z=-0.5)
    b.reset(origin_x=0.5, origin_y=0.25, origin_z=-0.5)
    assert np.array_equal(a.bins_y.numpy(), b.bins_y)


def test_cell_centres_equal_map_to_world(oracle, ref):
    kw = dict(camera_height=4, camera_width=4, map_height=12, map_width=10, map_depth=6,
              grid_resolution=0.05, origin_x=0.3, origin_y=-0.7, origin_z=0.9)
    a, b = ref.base.BaseProjectionLayer(**kw), oracle.OracleLayer(**kw)
    y, x, z = torch.meshgrid(torch.arange(12.), torch.arange(10.), torch.arange(6.), indexing="ij")
    world = a.map_to_world(torch.stack([x, y, z], dim=-1)).numpy()
    mx, my, mz = b.cell_centres()
    assert np.array_equal(world[..., 0], np.broadcast_to(mx[None, :, None], world.shape[:3]))
    assert np.array_equal(world[..., 1], np.broadcast_to(my[:, None, None], world.shape[:3]))
    assert np.array_equal(world[..., 2], np.broadcast_to(mz[None, None, :], world.shape[:3]))
