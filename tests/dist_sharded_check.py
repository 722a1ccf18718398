"""torchrun target (one rank per GPU, NCCL): frames of one scene sharded across the ranks with
update_batch_sharded, checked on every rank against the CPU oracle's sequential fusion -- once through the portable
all_gather exchange and once through the peer-memory exchange (rows read from the other ranks' allocations inside
the apply kernel).  With MASSB200_PEER_TEST_ONE_GPU=1 all ranks share cuda:0 and gloo is the control plane (NCCL does
not put two ranks on one device): the single-GPU test of the peer path.
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
    one_gpu = os.environ.get("MASSB200_PEER_TEST_ONE_GPU") == "1"
    if one_gpu:
        local = 0
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if one_gpu:
        dist.init_process_group("gloo")
    else:
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

    def verdict(layer, what):
        torch.cuda.synchronize()
        got = layer.data.cpu().numpy()
        occ = np.array_equal((got != 0).any(-1), (ref.data != 0).any(-1))
        err = np.abs(got.astype(np.float64) - ref.data)
        ok = bool((err <= 1e-5 * np.abs(ref.data)).all())
        flag = torch.tensor([int(ok and occ)], device="cpu" if one_gpu else dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if rank == 0:
            print("%s on %d ranks: %s (occupancy %s, max rel err %.2e)" % (
                what, world, "OK" if int(flag.item()) else "MISMATCH", occ,
                float((err / np.maximum(np.abs(ref.data), 1e-30)).max())), flush=True)
        return int(flag.item())

    good = 1
    if not one_gpu:
        layer = BaseProjectionLayer(exact=False, **kw).to(dev)
        layer.data.copy_(torch.from_numpy(start))
        sharded.update_batch_sharded(layer, mine)
        good &= verdict(layer, "all_gather exchange")
    # peer-memory exchange, twice through the same buffers (the second call checks that they come back empty)
    layer = BaseProjectionLayer(exact=False, **kw).to(dev)
    ex = sharded.PeerExchange(layer, capacity=60000)
    for rep in range(2):
        layer.data.copy_(torch.from_numpy(start))
        half = len(mine) // 2
        if half:                                       # a rank folds its chunk in two pieces (ring by ring)
            ex.partial.fold(layer, mine[:half])
        if rep == 0:                                   # pull all peers' rows, then apply locally
            sharded.update_batch_sharded(layer, mine[half:], exchange=ex)
        else:                                          # apply straight out of the owners' memory
            if mine[half:]:
                ex.partial.fold(layer, mine[half:])
            ex.combine(layer, pull=False)
        layer.check()
        good &= verdict(layer, "peer exchange (%s)" % ("pulled" if rep == 0 else "read in place"))
    ex.close()
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if good else 1)


if __name__ == "__main__":
    main()
