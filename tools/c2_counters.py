"""Prints the device counters of the batched pipeline after one C2 step (valid pixels, cells, segments,
runs, touched voxels).  Usage (GPU box): python tools/c2_counters.py [frames]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from mass_b200.nn.base_projection_layer import BaseProjectionLayer
from mass_b200.utils import synthetic

T = int(sys.argv[1]) if len(sys.argv) > 1 else 500
dev = torch.device("cuda:0")
walk = bench.make_walkthrough(T)
layer = BaseProjectionLayer(exact=False, **bench.C2, **synthetic.MAP_ORIGIN).to(dev)
depth = torch.from_numpy(walk["depth"]).to(dev)
probs = torch.from_numpy(walk["probs_low"]).to(dev).repeat_interleave(8, 1).repeat_interleave(8, 2).contiguous()
layer.update_batch(dict(position=walk["position"], yaw=walk["yaw"], elevation=walk["elevation"], depth=depth, features=probs))
torch.cuda.synchronize()
c = layer._last_ws[:64].view(torch.int32).cpu().tolist()
names = ["heads", "items", "cells", "segs", "runs", "error", "vox"]
print({n: c[i] for i, n in enumerate(names)}, "pixels", T * 224 * 224)
