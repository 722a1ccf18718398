"""Developer aid: per-phase cycle totals of k_brick_reduce_qw (library built with
`make EXTRA=-DMB_PHASE_TIMING`).  Runs the C2 bench workload for a few chunks."""
import ctypes, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from mass_b200 import _lib
from mass_b200.nn.base_projection_layer import BaseProjectionLayer
from mass_b200.utils import synthetic
T = int(sys.argv[1]) if len(sys.argv) > 1 else 128
dev = torch.device("cuda:0")
walk = bench.make_walkthrough(500)
layer = BaseProjectionLayer(exact=False, **bench.C2, **synthetic.MAP_ORIGIN).to(dev)
sl = slice(64, 64 + T)
depth = torch.from_numpy(walk["depth"][sl]).to(dev)
probs = torch.from_numpy(walk["probs_low"][sl]).to(dev).repeat_interleave(8, 1).repeat_interleave(8, 2).contiguous()
obs = dict(position=walk["position"][sl], yaw=walk["yaw"][sl], elevation=walk["elevation"][sl], depth=depth, features=probs)
L = ctypes.CDLL(_lib.SO_PATH)
out = (ctypes.c_ulonglong * 16)()
layer.update_batch(obs); L.mb_debug_phase_cycles(out, 1)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); layer.update_batch(obs); e1.record(); torch.cuda.synchronize()
L.mb_debug_phase_cycles(out, 1)
names = ["ticket+group setup", "chunk load+sync", "stage issue+weights+rank+sync", "prefix/scan/scatter", "cp.async wait+sync", "reduce+sync", "turn wait+sync", "apply+fence+publish"]
tot = sum(out[i] for i in range(8))
print("frames %d: %.3f ms; summed CTA cycles %.3g" % (T, e0.elapsed_time(e1), tot))
for i in range(8):
    print("  %-32s %5.1f%%" % (names[i], 100.0 * out[i] / max(tot, 1)))
