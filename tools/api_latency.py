"""Wall-clock latency of the drop-in per-frame call, layer.update(observation) with numpy observations exactly as
NavigationPolicy.process_observations passes them (H2D copies and host marshalling included), plus a cProfile of the
host side.  Usage (GPU box): python tools/api_latency.py"""
import cProfile
import os
import pstats
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import bench
from mass_b200.nn.applications.occupancy_projection_layer import OccupancyProjectionLayer
from mass_b200.nn.applications.semantic_projection_layer import SemanticProjectionLayer
from mass_b200.nn.base_projection_layer import BaseProjectionLayer
from mass_b200.utils import synthetic

dev = torch.device("cuda:0")
kw = dict(bench.C2, **synthetic.MAP_ORIGIN)
rays = synthetic.camera_rays(224, 224)
frames = [synthetic.boxroom_frame(t, 500, rays=rays) for t in range(40)]
for f in frames:
    f["semantic"] = f["features"].argmax(-1)[..., None].astype(np.int64)


def wall(layer, key_drop, n=30):
    obs = [{k: v for k, v in f.items() if k not in key_drop} for f in frames]
    for o in obs[:5]:
        layer.update(o)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for o in obs[5:5 + n]:
        layer.update(o)
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3


for exact in (True, False):
    print("exact=%s" % exact)
    print("  dense 54-class probabilities: %.3f ms per update()" % wall(BaseProjectionLayer(exact=exact, **kw).to(dev), ("semantic",)))
    print("  semantic ids                : %.3f ms per update()" % wall(SemanticProjectionLayer(exact=exact, **kw).to(dev), ("features",)))
    print("  occupancy                   : %.3f ms per update()" % wall(OccupancyProjectionLayer(exact=exact, **dict(kw, feature_size=1)).to(dev), ("features", "semantic")))

layer = SemanticProjectionLayer(exact=False, **kw).to(dev)
obs = [{k: v for k, v in f.items() if k != "features"} for f in frames]
for o in obs[:5]:
    layer.update(o)
pr = cProfile.Profile()
pr.enable()
for o in obs[5:35]:
    layer.update(o)
torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(14)
