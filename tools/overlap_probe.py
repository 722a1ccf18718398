"""Feasibility probe: how much of the instruction-bound front half of a batched update (voxelise, grouping, sort,
index, scalar pass) hides under the DRAM-bound feature pass of ANOTHER batch running on a second stream?

Two maps, two scratch buffers, two streams; each stream fuses 250 C2 frames per call.  Compared: both calls back
to back on one stream, and the two streams running concurrently with stream 2 started half a call late.
Run on a B200:  python tools/overlap_probe.py
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from bench import C2, make_walkthrough          # noqa: E402
from mass_b200 import _lib                      # noqa: E402
from mass_b200.nn.base_projection_layer import BaseProjectionLayer   # noqa: E402
from mass_b200.utils import synthetic           # noqa: E402


def main():
    dev = torch.device("cuda", 0)
    T = 500
    walk = make_walkthrough(T)
    depth = torch.from_numpy(walk["depth"]).to(dev)
    probs = torch.from_numpy(walk["probs_low"]).to(dev).repeat_interleave(8, dim=1).repeat_interleave(8, dim=2).contiguous()
    halves, layers = [], []
    for h in range(2):
        sl = slice(h * T // 2, (h + 1) * T // 2)
        layer = BaseProjectionLayer(exact=False, **C2, **synthetic.MAP_ORIGIN).to(dev)
        layer._ws = _lib.Workspace()
        obs = dict(position=walk["position"][sl], yaw=walk["yaw"][sl], elevation=walk["elevation"][sl],
                   depth=depth[sl], features=probs[sl])
        halves.append(layer.prepare_batch(obs))
        layers.append(layer)
    reps = 10
    streams = [torch.cuda.Stream(dev), torch.cuda.Stream(dev)]

    def run(concurrent, delay_ms):
        torch.cuda.synchronize()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0, s1 = streams
        ev0.record(torch.cuda.current_stream())
        s0.wait_event(ev0), s1.wait_event(ev0)
        if concurrent:
            with torch.cuda.stream(s1):
                torch.cuda._sleep(int(delay_ms * 1.9e6))
            for _ in range(reps):
                with torch.cuda.stream(s0):
                    layers[0].update_prepared(halves[0])
                with torch.cuda.stream(s1):
                    layers[1].update_prepared(halves[1])
        else:
            with torch.cuda.stream(s0):
                for _ in range(reps):
                    layers[0].update_prepared(halves[0])
                    layers[1].update_prepared(halves[1])
        done0, done1 = torch.cuda.Event(), torch.cuda.Event()
        done0.record(s0), done1.record(s1)
        torch.cuda.current_stream().wait_event(done0), torch.cuda.current_stream().wait_event(done1)
        ev1.record(torch.cuda.current_stream())
        torch.cuda.synchronize()
        return ev0.elapsed_time(ev1) / reps

    def graph_time(concurrent):
        g = torch.cuda.CUDAGraph()
        s0, s1 = streams
        torch.cuda.synchronize()
        with torch.cuda.graph(g, stream=s0, capture_error_mode="thread_local"):
            if concurrent:
                fork = torch.cuda.Event()
                fork.record(s0)
                s1.wait_event(fork)
                layers[0].update_prepared(halves[0])
                with torch.cuda.stream(s1):
                    layers[1].update_prepared(halves[1])
                    join = torch.cuda.Event()
                    join.record(s1)
                s0.wait_event(join)
            else:
                layers[0].update_prepared(halves[0])
                layers[1].update_prepared(halves[1])
        for _ in range(3):
            g.replay()
        torch.cuda.synchronize()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        for _ in range(reps):
            g.replay()
        ev1.record()
        torch.cuda.synchronize()
        return ev0.elapsed_time(ev1) / reps

    def pipelined(nbatches=20):
        """steady state: half batches alternate between the two streams, each stream's calls back to back"""
        torch.cuda.synchronize()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0, s1 = streams
        ev0.record(torch.cuda.current_stream())
        s0.wait_event(ev0), s1.wait_event(ev0)
        for i in range(nbatches):
            with torch.cuda.stream(streams[i % 2]):
                layers[i % 2].update_prepared(halves[i % 2])
        d0, d1 = torch.cuda.Event(), torch.cuda.Event()
        d0.record(s0), d1.record(s1)
        torch.cuda.current_stream().wait_event(d0), torch.cuda.current_stream().wait_event(d1)
        ev1.record(torch.cuda.current_stream())
        torch.cuda.synchronize()
        return ev0.elapsed_time(ev1) / (nbatches / 2)

    for _ in range(2):
        run(False, 0)
    pipelined()
    print("two streams, 20 half batches alternating: %.3f ms per 500 frames" % pipelined())
    print("graph, 2 x 250 frames back to back:      %.3f ms per 500 frames" % graph_time(False))
    print("graph, 2 x 250 frames on two branches:   %.3f ms per 500 frames" % graph_time(True))
    print("one stream, 2 x 250 frames back to back: %.3f ms per 500 frames" % run(False, 0))
    for delay in (0.0, 0.4, 0.8, 1.2):
        run(True, delay)
        print("two streams, second delayed %.1f ms:        %.3f ms per 500 frames" % (delay, run(True, delay)))


if __name__ == "__main__":
    main()
