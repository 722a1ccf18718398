"""BASELINE config 4 in the small: frames of ONE large scene (720x1280 depth + 54-class probabilities, 0.02 m grid)
split into contiguous chunks across the ranks, partial maps combined by the ordered affine exchange over NCCL
(mass_b200/nn/sharded.py).  Prints per-phase times and checks the result against rank-local sequential fusion.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/c4_sharded.py [frames_per_rank]"""
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    from mass_b200.nn.base_projection_layer import BaseProjectionLayer
    from mass_b200.nn import sharded
    from mass_b200.utils import synthetic

    per_rank = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    H, W, F = 720, 1280, 54
    T = per_rank * world
    kw = dict(camera_height=H, camera_width=W, vertical_fov=90.0, map_height=640, map_width=640, map_depth=160,
              feature_size=F, grid_resolution=0.02, interpolation_weight=0.5, **synthetic.MAP_ORIGIN)
    rays = synthetic.camera_rays(H, W)
    lo = rank * per_rank

    def frames(t0, t1):
        pos, yaw, elev, depth = [], [], [], []
        for t in range(t0, t1):
            p, y, e = synthetic.boxroom_pose(t, 512)          # a slow orbit: consecutive frames overlap heavily
            d, _ = synthetic.render_depth(rays, p, y, e)
            pos.append(p), yaw.append(y), elev.append(e), depth.append(d[..., None])
        g = torch.Generator(device=dev).manual_seed(1000 + t0)
        low = torch.softmax(4 * torch.randn(t1 - t0, H // 8, W // 8, F, device=dev, generator=g), dim=-1)
        probs = low.repeat_interleave(8, 1).repeat_interleave(8, 2).contiguous()
        return dict(position=np.stack(pos), yaw=np.array(yaw, np.float32), elevation=np.array(elev, np.float32),
                    depth=torch.from_numpy(np.stack(depth)).to(dev), features=probs)

    mine = frames(lo, lo + per_rank)
    layer = BaseProjectionLayer(exact=False, **kw).to(dev)
    partial = sharded.PartialMap(layer)
    torch.cuda.synchronize()

    def step():
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        ev[0].record()
        idx, a, b = sharded.fold_frames(layer, mine, partial)
        ev[1].record()
        parts = sharded.exchange_partials(idx, a, b) if world > 1 else [(idx, a, b)]
        ev[2].record()
        for gi, ga, gb in parts:
            sharded.apply_partial(layer, gi, ga, gb)
        ev[3].record()
        torch.cuda.synchronize()
        return [ev[i].elapsed_time(ev[i + 1]) for i in range(3)], sum(int(p[0].numel()) for p in parts)

    step()                                                       # warm-up (allocations, NCCL channels)
    layer.data.zero_()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    times, touched = step()
    if world > 1:
        dist.barrier()
    wall = time.perf_counter() - t0
    tt = torch.tensor(times + [wall * 1e3], device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    # check: rank 0 fuses all T frames sequentially into a second map and compares
    ok = True
    if rank == 0:
        ref = BaseProjectionLayer(exact=False, **kw).to(dev)
        for g in range(world):
            ref.update_batch(frames(g * per_rank, (g + 1) * per_rank))
        torch.cuda.synchronize()
        occ_equal, rel_ok = True, True
        for y in range(0, ref.data.shape[0], 32):                # slab by slab: no map-sized temporaries
            r, l = ref.data[y:y + 32], layer.data[y:y + 32]
            occ_equal = occ_equal and bool(torch.equal((r != 0).any(-1), (l != 0).any(-1)))
            rel_ok = rel_ok and bool(((r - l).abs() <= 2e-5 * r.abs()).all())
        ok = occ_equal and rel_ok
        fold, exch, app, w = [float(x) for x in tt.tolist()]
        print("c4-small: %d ranks x %d frames of %dx%dx%d into a %dx%dx%d map at 0.02 m" % (world, per_rank, H, W, F, 640, 640, 160))
        print("  fold %.1f ms, exchange %.1f ms, ordered apply %.1f ms, wall %.1f ms -> %.0f frames/s; "
              "%d partial rows exchanged (%.1f MB); sharded == sequential: occupancy %s, values within 2e-5: %s"
              % (fold, exch, app, w, T / (w * 1e-3), touched, touched * (F + 3) * 4 / 1e6, occ_equal, rel_ok))
    if world > 1:
        dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
