"""Cosine similarity + best match: the tcgen05 path against the SIMT kernel on large instance matrices (the sizes at
which the instance-by-feature matrix is a real dense contraction).  Prints ms per call, the tensor-core path's
TFLOP/s on the 3 x TF32 products of its two passes, and checks that both paths return the same indices.
    python tools/match_bench.py [n ...]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mass_b200.utils import instances  # noqa: E402


def timed(fn, reps):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, out


def main():
    dev = torch.device("cuda:0")
    sizes = [int(x) for x in sys.argv[1:]] or [1024, 4096, 16384]
    d = 256
    for n in sizes:
        g = torch.Generator(device=dev).manual_seed(n)
        a = torch.randn(n, d, device=dev, generator=g)
        b = torch.cat([a[torch.randperm(n, device=dev, generator=g)[:n // 2]] * 1.3 + 0.05 * torch.randn(n // 2, d, device=dev, generator=g),
                       torch.randn(n - n // 2, d, device=dev, generator=g)])
        ms_tc, (best_tc, sim_tc) = timed(lambda: instances.cosine_best_match(a, b, tensor_cores=True), 5)
        reps = 3 if n <= 4096 else 1
        ms_sm, (best_sm, sim_sm) = timed(lambda: instances.cosine_best_match(a, b, tensor_cores=False), reps)
        same = bool(torch.equal(best_tc, best_sm))
        flops = 2.0 * n * n * d * 3 * 2                     # three TF32 products per pass, two passes
        print("n = m = %6d, d = %d: tcgen05 %8.3f ms (%.1f TFLOP/s on its 3xTF32 products), SIMT float64 %9.3f ms, "
              "speed-up %.1fx, same indices: %s" % (n, d, ms_tc, flops / (ms_tc * 1e-3) / 1e12, ms_sm, ms_sm / ms_tc, same))


if __name__ == "__main__":
    main()
