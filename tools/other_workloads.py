"""Times the batched update on the other layer shapes of an episode (GPU box):
   semantic map fed class ids (one-hot without materialising it), 256-d feature map at a quarter of the camera
   resolution, occupancy map (F = 1), and the per-frame exact mode.  Usage: python tools/other_workloads.py [frames]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import bench
from mass_b200.nn.base_projection_layer import BaseProjectionLayer
from mass_b200.nn.applications.semantic_projection_layer import SemanticProjectionLayer
from mass_b200.utils import synthetic

T = int(sys.argv[1]) if len(sys.argv) > 1 else 500
dev = torch.device("cuda:0")
walk = bench.make_walkthrough(T)
depth = torch.from_numpy(walk["depth"]).to(dev)
base = dict(position=walk["position"], yaw=walk["yaw"], elevation=walk["elevation"])


def timeit(fn, reps=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

def stages(fn):
    """Stage timers of one call (voxelise, sort, index, scalar pass, accumulate, apply), ms."""
    import ctypes
    from mass_b200 import _lib
    L = _lib.lib()
    L.mb_profile_stages(1)
    fn()
    buf = (ctypes.c_float * 8)()
    n = L.mb_profile_read(buf, 8)
    L.mb_profile_stages(0)
    return "stages ms: " + " ".join("%.2f" % buf[i] for i in range(n))


kw = dict(bench.C2, **synthetic.MAP_ORIGIN)
# semantic ids
ids = torch.from_numpy(walk["probs_low"]).to(dev).argmax(-1).repeat_interleave(8, 1).repeat_interleave(8, 2).contiguous()
sem = SemanticProjectionLayer(exact=False, **kw).to(dev)
prep = sem.prepare_batch(dict(base, depth=depth, class_ids=ids))
ms = timeit(lambda: sem.update_prepared(prep))
print("semantic map from class ids, %d frames 224x224: %.2f ms  (%.0f frames/s)" % (T, ms, T / ms * 1e3))
print("   ", stages(lambda: sem.update_prepared(prep)))
del sem
# 256-d feature map at 56x56
kw256 = dict(kw, camera_height=56, camera_width=56, feature_size=256)
feat = torch.rand(T, 56, 56, 256, device=dev)
res = BaseProjectionLayer(exact=False, **kw256).to(dev)
prep = res.prepare_batch(dict(base, depth=depth[:, 2::4, 2::4].contiguous(), features=feat))
ms = timeit(lambda: res.update_prepared(prep))
gb = T * 56 * 56 * 256 * 4 / 1e9
print("256-d feature map, %d frames 56x56: %.2f ms  (%.0f frames/s, %.0f GB/s of feature rows)" % (T, ms, T / ms * 1e3, gb / ms * 1e3))
del res, feat
# occupancy
kw1 = dict(kw, feature_size=1)
occ = BaseProjectionLayer(exact=False, **kw1).to(dev)
prep = occ.prepare_batch(dict(base, depth=depth, features=torch.ones(T, 224, 224, 1, device=dev)))
ms = timeit(lambda: occ.update_prepared(prep))
print("occupancy map (F = 1), %d frames 224x224: %.2f ms  (%.0f frames/s)" % (T, ms, T / ms * 1e3))
print("   ", stages(lambda: occ.update_prepared(prep)))
del occ
# exact per-frame mode, dense probabilities
probs = torch.from_numpy(walk["probs_low"][:32]).to(dev).repeat_interleave(8, 1).repeat_interleave(8, 2).contiguous()
ex = BaseProjectionLayer(exact=True, **kw).to(dev)
prep = ex.prepare_batch(dict(position=walk["position"][:32], yaw=walk["yaw"][:32], elevation=walk["elevation"][:32],
                             depth=depth[:32], features=probs))
ms = timeit(lambda: ex.update_prepared(prep))
print("exact mode (bitwise), 32 frames 224x224x54: %.2f ms  (%.0f frames/s)" % (ms, 32 / ms * 1e3))
# one frame per call (the reference's call pattern: NavigationPolicy.process_observations), batched arithmetic
fast = BaseProjectionLayer(exact=False, **kw).to(dev)
preps = [fast.prepare_batch(dict(position=walk["position"][t:t + 1], yaw=walk["yaw"][t:t + 1], elevation=walk["elevation"][t:t + 1],
                                 depth=depth[t:t + 1], features=probs[t:t + 1])) for t in range(32)]
ms = timeit(lambda: [fast.update_prepared(p) for p in preps])
print("one frame per call, batched arithmetic, 32 calls 224x224x54: %.2f ms  (%.0f frames/s, %.0f us per call)" % (ms, 32 / ms * 1e3, ms / 32 * 1e3))
