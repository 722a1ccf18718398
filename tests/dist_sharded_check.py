"""torchrun target (one rank per GPU, NCCL): frames of one scene sharded across the ranks with
update_batch_sharded, checked on every rank against the CPU oracle's sequential fusion.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tests/dist_sharded_check.py"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    from mass_b200.nn.base_projection_layer import BaseProjectionLayer
    from mass_b200.nn import sharded
    from mass_b200.utils import synthetic
    from oracle import oracle

    H = W = 64
    T = 4 * world + 1                                   # uneven split on purpose
    kw = dict(camera_height=H, camera_width=W, vertical_fov=90.0, map_height=96, map_width=96, map_depth=32,
              feature_size=54, grid_resolution=0.2, interpolation_weight=0.5, **synthetic.MAP_ORIGIN)
    frames = [synthetic.boxroom_frame(t, T, height=H, width=W, feature_size=54) for t in range(T)]
    ref = oracle.OracleLayer(**kw)
    start = np.random.default_rng(3).random(ref.data.shape).astype(np.float32) * (np.random.default_rng(4).random(ref.data.shape[:3])[..., None] < 0.3)
    ref.data[...] = start
    for f in frames:
        ref.update(f)
    bounds = np.linspace(0, T, world + 1).astype(int)
    mine = frames[bounds[rank]:bounds[rank + 1]]
    layer = BaseProjectionLayer(exact=False, **kw).to(dev)
    layer.data.copy_(torch.from_numpy(start))
    sharded.update_batch_sharded(layer, mine)
    torch.cuda.synchronize()
    got = layer.data.cpu().numpy()
    occ = np.array_equal((got != 0).any(-1), (ref.data != 0).any(-1))
    err = np.abs(got.astype(np.float64) - ref.data)
    ok = bool((err <= 1e-5 * np.abs(ref.data)).all())
    flag = torch.tensor([int(ok and occ)], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("sharded fusion on %d ranks: %s (occupancy %s, max rel err %.2e)" % (
            world, "OK" if int(flag.item()) else "MISMATCH", occ,
            float((err / np.maximum(np.abs(ref.data), 1e-30)).max())))
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) else 1)


if __name__ == "__main__":
    main()
