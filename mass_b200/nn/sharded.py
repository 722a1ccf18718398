"""Frames of ONE scene sharded across GPUs (BASELINE config 4).

The reference is single-GPU; what it fixes is the per-voxel update, which is affine in the old row and does
not commute across frames (/root/reference/mass/utils/projection.py:335-351).  Frames are
therefore split into CONTIGUOUS chunks in time order, rank g taking chunk g.  Every rank folds its chunk,
from the identity, into a sparse partial {voxel index, A, B rows}: the chunk acts on any map M as
M[v] <- A[v] * M[v] + B[v].  The partials are all-gathered (NCCL over NVLink; the only exchange step of
the path) and every rank applies them to its replica of the map in rank order, which is exactly the
sequential semantics -- a plain sum of partial grids would be wrong.  The exchanged volume is
(F + 2) * 4 bytes per touched voxel, not the dense map.
"""
import torch
import torch.distributed as dist

from mass_b200 import _lib


class PartialMap:
    """Dense scratch pair a rank folds its frames into, re-used across calls."""

    def __init__(self, layer):
        data = layer.data
        self.voxels = data.shape[0] * data.shape[1] * data.shape[2]
        self.features = data.shape[3]
        self.b = torch.zeros_like(data)
        self.a = torch.full((self.voxels,), 2.0, dtype=torch.float32, device=data.device)   # 2.0 = untouched

    def extract(self):
        """Sparse partial (idx int64 [n], a [n], b [n, F]) of everything folded since the last reset."""
        idx = (self.a != 2.0).nonzero(as_tuple=False).reshape(-1)
        return idx, self.a[idx], self.b.view(self.voxels, self.features)[idx]

    def reset(self, idx):
        self.a[idx] = 2.0
        self.b.view(self.voxels, self.features)[idx] = 0.0


def fold_frames(layer, observations, partial=None):
    """Fold `observations` (a contiguous run of frames) into a sparse partial without touching layer.data."""
    partial = partial or PartialMap(layer)
    layer.update_batch(observations, fold=(partial.b, partial.a))
    idx, a, b = partial.extract()
    partial.reset(idx)
    return idx, a, b


def apply_partial(layer, idx, a, b):
    """layer.data[idx] = a * layer.data[idx] + b, in place (one rank's partial)."""
    data = layer.data
    device = _lib.require_cuda(data.device)
    n = int(idx.numel())
    if n == 0:
        return
    _lib.check(_lib.lib().mb_affine_apply_rows(
        _lib.stream_ptr(device), _lib.ptr(data), int(data.shape[3]), _lib.ptr(idx.to(torch.int64).contiguous()),
        _lib.ptr(a.to(torch.float32).contiguous()), _lib.ptr(b.to(torch.float32).contiguous()), n))
    layer.mark_dirty()            # a raw-pointer write: results memoised against map_state() are stale now


def exchange_partials(idx, a, b, group=None):
    """All-gathers every rank's sparse partial.  Returns a list, in rank (= time) order, of (idx, a, b).
    Works on any backend (NCCL on GPUs; gloo in the CPU tests): two collectives, sizes then payload."""
    world = dist.get_world_size(group)
    feat = int(b.shape[1]) if b.dim() == 2 else 0
    n = torch.tensor([idx.numel()], dtype=torch.int64, device=idx.device)
    sizes = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(sizes, n, group=group)
    sizes = [int(s.item()) for s in sizes]
    nmax = max(sizes)
    # one payload per rank: [nmax, F + 3] float32 = {index low 24 bits, index high bits, a, b...}
    # (indices < 2^48 survive the trip through float32 exactly as two 24-bit halves)
    pay = torch.zeros(nmax, feat + 3, dtype=torch.float32, device=idx.device)
    k = idx.numel()
    if k:
        i64 = idx.to(torch.int64)
        pay[:k, 0] = (i64 & 0xFFFFFF).to(torch.float32)
        pay[:k, 1] = (i64 >> 24).to(torch.float32)
        pay[:k, 2] = a
        pay[:k, 3:] = b
    gathered = [torch.empty_like(pay) for _ in range(world)]
    dist.all_gather(gathered, pay, group=group)
    out = []
    for g, buf in enumerate(gathered):
        m = sizes[g]
        gi = buf[:m, 0].to(torch.int64) | (buf[:m, 1].to(torch.int64) << 24)
        out.append((gi, buf[:m, 2].contiguous(), buf[:m, 3:].contiguous()))
    return out


def update_batch_sharded(layer, local_observations, group=None, partial=None, apply_fn=None):
    """Collective call: every rank passes ITS contiguous chunk of the scene's frames (rank order = time order;
    a rank may pass None or an empty chunk).  On return every rank's layer.data holds the map after all frames,
    equal to sequential fusion up to fp32 re-association (occupancy identical)."""
    if local_observations is not None and _num_frames(local_observations) > 0:
        idx, a, b = fold_frames(layer, local_observations, partial)
    else:
        dev, feat = layer.data.device, layer.data.shape[3]
        idx = torch.zeros(0, dtype=torch.int64, device=dev)
        a = torch.zeros(0, dtype=torch.float32, device=dev)
        b = torch.zeros(0, feat, dtype=torch.float32, device=dev)
    apply_fn = apply_fn or apply_partial
    for gi, ga, gb in exchange_partials(idx, a, b, group):
        apply_fn(layer, gi, ga, gb)
    return layer


def _num_frames(observations):
    if isinstance(observations, (list, tuple)):
        return len(observations)
    return int(torch.as_tensor(observations["yaw"]).reshape(-1).shape[0])
