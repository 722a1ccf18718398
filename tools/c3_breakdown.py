"""Where the mapping half of a C3 episode pair goes: device time (CUDA events) and host time (wall clock, no
synchronisation in between) of every reset and update_batch of the five maps.  Run on a B200: python tools/c3_breakdown.py"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench_configs as bc                      # noqa: E402
from mass_b200.utils import synthetic           # noqa: E402


def main():
    dev = torch.device("cuda", 0)
    feat_table = torch.rand(bc.C3_OBJECTS + 1, 256, device=dev, generator=torch.Generator(device=dev).manual_seed(7))
    scenes = [bc.c3_scene(T, shifted, feat_table, dev) for shifted, T in zip((False, True), bc.C3_FRAMES)]
    occ, sems, ress = bc.c3_layers(dev)
    origin = {k: synthetic.MAP_ORIGIN[k] for k in ("origin_y", "origin_x", "origin_z")}
    steps = []
    for name, layer in (("occ", occ), ("sem0", sems[0]), ("sem1", sems[1]), ("res0", ress[0]), ("res1", ress[1])):
        steps.append(("reset " + name, lambda layer=layer: layer.reset(**origin)))
    for i, obs in enumerate(scenes):
        if i == 1:
            steps.append(("reset occ", lambda: occ.reset(**origin)))
        steps.append(("occ.update_batch[%d]" % i, lambda obs=obs: occ.update_batch(obs)))
        steps.append(("sem%d.update_batch" % i, lambda obs=obs, i=i: sems[i].update_batch(obs)))
        steps.append(("res%d.update_batch" % i, lambda obs=obs, i=i: ress[i].update_batch(obs)))
    for _ in range(2):                           # warm-up
        for _, fn in steps:
            fn()
    torch.cuda.synchronize()
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(len(steps) + 1)]
    host = []
    t_all = time.perf_counter()
    evs[0].record()
    for k, (_, fn) in enumerate(steps):
        t0 = time.perf_counter()
        fn()
        host.append((time.perf_counter() - t0) * 1e3)
        evs[k + 1].record()
    t_host = (time.perf_counter() - t_all) * 1e3
    torch.cuda.synchronize()
    t_wall = (time.perf_counter() - t_all) * 1e3
    print("%-24s %10s %10s" % ("step", "device ms", "host ms"))
    for k, (name, _) in enumerate(steps):
        print("%-24s %10.3f %10.3f" % (name, evs[k].elapsed_time(evs[k + 1]), host[k]))
    print("device total %.2f ms, host enqueue total %.2f ms, wall %.2f ms" % (evs[0].elapsed_time(evs[-1]), t_host, t_wall))
    # stage timers (voxelise, sort, index, scalar pass, accumulate, apply) and device counters of the LAST chunk of each update
    import ctypes
    from mass_b200 import _lib
    L = _lib.lib()
    for name, layer, obs in (("occ[1]", occ, scenes[1]), ("sem1", sems[1], scenes[1]), ("res1", ress[1], scenes[1])):
        L.mb_profile_stages(1)
        layer.update_batch(obs)
        buf = (ctypes.c_float * 8)()
        n = L.mb_profile_read(buf, 8)
        L.mb_profile_stages(0)
        print("%-8s stages ms: %s   counters %s" % (name, " ".join("%.3f" % buf[i] for i in range(n)), layer.counters()))


if __name__ == "__main__":
    main()
