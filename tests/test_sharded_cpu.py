"""Host logic of the frame-sharded scene fusion (mass_b200/nn/sharded.py) on CPU, world_size 2, gloo.

The partials come from the CPU oracle (a chunk's action on the map is affine per voxel: running it from an
all-zero and from an all-one map gives B and A + B), travel through exchange_partials, and are applied in rank
order by a numpy stand-in for the CUDA apply kernel.  The result must equal the oracle's sequential fusion."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

KW = dict(camera_height=16, camera_width=20, vertical_fov=90.0, map_height=30, map_width=28, map_depth=12,
          feature_size=5, grid_resolution=0.15, interpolation_weight=0.5, origin_z=0.3)
T = 6


def _frames():
    rng = np.random.default_rng(42)
    H, W, F = KW["camera_height"], KW["camera_width"], KW["feature_size"]
    return dict(position=rng.uniform(-.4, .4, (T, 3)).astype(np.float32), yaw=rng.uniform(-3, 3, T).astype(np.float32),
                elevation=rng.uniform(-.5, .5, T).astype(np.float32),
                depth=rng.uniform(0.3, 2.0, (T, H, W, 1)).astype(np.float32),
                features=rng.random((T, H, W, F)).astype(np.float32))


def _run(oracle, frames, ts, start):
    layer = oracle.OracleLayer(**KW)
    layer.data[...] = start
    for t in ts:
        layer.update({k: v[t] for k, v in frames.items()})
    return layer.data.copy()


def _partial(oracle, frames, ts):
    F = KW["feature_size"]
    b = _run(oracle, frames, ts, 0.0).reshape(-1, F)
    ab = _run(oracle, frames, ts, 1.0).reshape(-1, F)
    a = (ab - b)[:, 0]
    idx = np.flatnonzero(a != 1.0)
    return torch.from_numpy(idx.astype(np.int64)), torch.from_numpy(a[idx].copy()), torch.from_numpy(b[idx].copy())


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import sys
        sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
        from mass_b200.nn import sharded
        from oracle import oracle
        frames = _frames()
        chunks = [range(0, 4), range(4, 6)]               # contiguous, unequal
        idx, a, b = _partial(oracle, frames, chunks[rank])
        start = np.random.default_rng(7).random((KW["map_height"], KW["map_width"], KW["map_depth"],
                                                 KW["feature_size"])).astype(np.float32)
        data = start.copy().reshape(-1, KW["feature_size"])
        order = []
        for g, (gi, ga, gb) in enumerate(sharded.exchange_partials(idx, a, b)):
            gi, ga, gb = gi.numpy(), ga.numpy(), gb.numpy()
            order.append(int(gi.size))
            data[gi] = ga[:, None] * data[gi] + gb
        ref = _run(oracle, frames, range(T), start).reshape(-1, KW["feature_size"])
        err = np.abs(data.astype(np.float64) - ref)
        ok = bool((err <= 2e-5 * np.abs(ref) + 1e-12).all())
        occ = bool(np.array_equal(data != 0, ref != 0))
        # a rank with no frames takes part in the collective with an empty partial
        empty = sharded.exchange_partials(torch.zeros(0, dtype=torch.int64), torch.zeros(0), torch.zeros(0, 5)) \
            if rank == 1 else sharded.exchange_partials(idx, a, b)
        out[rank] = (ok, occ, order, [int(e[0].numel()) for e in empty], float((err / np.maximum(np.abs(ref), 1e-30)).max()))
    finally:
        dist.destroy_process_group()


def test_ordered_affine_combine_gloo_world2():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    assert len(out) == 2
    for rank in (0, 1):
        ok, occ, order, empty, worst = out[rank]
        assert ok and occ, (rank, worst)
        assert empty[1] == 0 and empty[0] == order[0]
    assert out[0][2] == out[1][2]            # both ranks saw the partials in the same (rank) order


def test_a_plain_sum_of_partials_is_wrong(oracle):
    """the combine must be the ordered affine composition, not a reduction by sum."""
    frames = _frames()
    F = KW["feature_size"]
    b0 = _run(oracle, frames, range(0, 4), 0.0)
    b1 = _run(oracle, frames, range(4, 6), 0.0)
    ref = _run(oracle, frames, range(T), 0.0)
    both = (b0 != 0).any(-1) & (b1 != 0).any(-1)
    assert both.any()
    rel = np.abs((b0 + b1)[both] - ref[both]) / np.maximum(np.abs(ref[both]), 1e-30)
    assert rel.max() > 1e-2
