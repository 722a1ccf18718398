"""Frames of ONE scene sharded across GPUs (BASELINE config 4).

The reference is single-GPU; what it fixes is the per-voxel update, which is affine in the old row and does
not commute across frames (/root/reference/mass/utils/projection.py:335-351).  Frames are
therefore split into CONTIGUOUS chunks in time order, rank g taking chunk g.  Every rank folds its chunk,
from the identity, into a SPARSE partial {voxel index, a, b rows}: the chunk acts on any map M as
M[v] <- a[v] * M[v] + b[v].  Every rank then applies all partials to its replica of the map in rank (= time)
order, which is exactly the sequential semantics -- a plain sum of partial grids would be wrong.

No second dense map anywhere: a partial is `capacity` rows of (8 + 4 + 4 F) bytes plus a 4-byte-per-voxel slot
table, folded by the library's own kernels (`mb_layer_fold_sparse`: the rows come straight off the touched-voxel
list of the batched pipeline).  Two ways to the other ranks' partials:

  * `PeerExchange` (NVLink / NVSwitch, the product path): every rank's partial buffer is its own allocation,
    mapped by all peers through CUDA IPC; `mb_affine_apply_partial` reads the rows of rank g straight out of rank
    g's memory while applying them -- transfer and application are ONE kernel per partial, the row count is read
    on the device, and the only host-visible steps are two stream-ordered barriers;
  * `exchange_partials` (any torch.distributed backend: NCCL all_gather on GPUs, gloo in the CPU tests): sizes,
    then index / a / b payloads padded to the largest rank.
"""
import ctypes

import torch
import torch.distributed as dist

from mass_b200 import _lib


class _RawCuda:
    """A device allocation the library made (cudaMalloc), dressed for torch.as_tensor."""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (int(nbytes),), "typestr": "|u1", "data": (int(ptr), False),
                                         "version": 2}


class SparsePartial:
    """The action of a contiguous run of frames on any map, as `capacity` rows {index, a, b[F]} plus the slot table
    that finds a voxel's row while chunks are folded in.  `peer=True` puts the rows in an allocation other
    processes of the box can map (`handle()` / PeerExchange)."""

    def __init__(self, layer, capacity, peer=False):
        data = layer.data
        self.device = _lib.require_cuda(data.device)
        self.voxels = data.shape[0] * data.shape[1] * data.shape[2]
        self.features = int(data.shape[3])
        self.capacity = int(capacity)
        L = _lib.lib()
        self.nbytes = int(L.mb_partial_buffer_bytes(self.capacity, self.features))
        off = (ctypes.c_size_t * 4)()
        _lib.check(L.mb_partial_buffer_layout(self.capacity, self.features, off))
        self._peer_ptr = None
        if peer:
            with torch.cuda.device(self.device):
                p = ctypes.c_void_p(0)
                _lib.check(L.mb_peer_alloc(self.nbytes, ctypes.byref(p)))
            self._peer_ptr = p.value
            self.buffer = torch.as_tensor(_RawCuda(p.value, self.nbytes), device=self.device)
        else:
            self.buffer = torch.zeros(self.nbytes, dtype=torch.uint8, device=self.device)
        self.buffer_ptr = ctypes.c_void_p(self.buffer.data_ptr())
        cap, F = self.capacity, self.features
        self.count_view = self.buffer[off[0]:off[0] + 4].view(torch.int32)
        self.index = self.buffer[off[1]:off[1] + 8 * cap].view(torch.int64)
        self.a = self.buffer[off[2]:off[2] + 4 * cap].view(torch.float32)
        self.b = self.buffer[off[3]:off[3] + 4 * cap * F].view(torch.float32).view(cap, F)
        self.slot_table = torch.empty(self.voxels, dtype=torch.int32, device=self.device)
        _lib.check(L.mb_partial_reset(_lib.stream_ptr(self.device), _lib.ptr(self.slot_table), self.voxels, self.buffer_ptr))

    def fold(self, layer, observations):
        """Composes `observations` (the NEXT frames of this rank's chunk, in order) onto the partial; layer.data is
        not touched."""
        layer.update_batch(observations, fold=self)
        return self

    def count(self):
        """Rows in use (synchronises)."""
        return min(int(self.count_view.item()), self.capacity)

    def rows(self):
        """(index int64 [n], a [n], b [n, F]) views of the rows in use (synchronises)."""
        n = self.count()
        return self.index[:n], self.a[:n], self.b[:n]

    def clear(self):
        """Back to the identity: the slot-table entries of the rows in use are reset (no map-sized memset)."""
        _lib.check(_lib.lib().mb_partial_clear(_lib.stream_ptr(self.device), _lib.ptr(self.slot_table), self.buffer_ptr,
                                               self.capacity, self.features))

    def handle(self):
        """64-byte CUDA IPC handle of the row buffer (peer=True only)."""
        if self._peer_ptr is None:
            raise RuntimeError("this partial was not allocated as peer memory")
        h = (ctypes.c_ubyte * 64)()
        _lib.check(_lib.lib().mb_peer_export(ctypes.c_void_p(self._peer_ptr), h))
        return bytes(h)

    def __del__(self):
        try:
            if self._peer_ptr is not None:
                self.buffer = self.index = self.a = self.b = self.count_view = None
                _lib.lib().mb_peer_free(ctypes.c_void_p(self._peer_ptr))
        except Exception:
            pass


def apply_partial_buffer(layer, buffer_ptr, capacity):
    """layer.data[index[i]] = a[i] * layer.data[index[i]] + b[i] for the rows of the partial at `buffer_ptr` -- a
    local SparsePartial's buffer or a PEER's (then the rows cross NVLink inside this kernel).  The row count is read on
    the device: no host round trip."""
    data = layer.data
    device = _lib.require_cuda(data.device)
    _lib.check(_lib.lib().mb_affine_apply_partial(_lib.stream_ptr(device), _lib.ptr(data), int(data.shape[3]),
                                                  buffer_ptr, int(capacity)))
    layer.mark_dirty()


def _rank_barrier(group, device):
    """All ranks' work enqueued so far is complete before any rank's later work starts.  NCCL: a one-element
    all_reduce, ordered on the stream (the host does not wait).  Other backends: synchronise + host barrier."""
    if dist.get_backend(group) == "nccl":
        dist.all_reduce(torch.zeros(1, device=device), group=group)
    else:
        torch.cuda.synchronize(device)
        dist.barrier(group=group)


class PeerExchange:
    """Collective set-up, once: every rank allocates its partial as peer memory, the IPC handles go round
    (all_gather_object), every rank maps every other rank's buffer.  `combine(layer)` then applies all ranks'
    partials, in rank order, reading each one where it lives."""

    def __init__(self, layer, capacity, group=None):
        self.group = group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.partial = SparsePartial(layer, capacity, peer=True)
        self.capacity = int(capacity)
        self._slots = self._stage = None
        handles = [None] * self.world
        dist.all_gather_object(handles, (self.partial.handle(), self.capacity), group=group)
        L = _lib.lib()
        self.ptrs, self._opened = [], []
        with torch.cuda.device(self.partial.device):
            for g, (h, cap) in enumerate(handles):
                if cap != self.capacity:
                    raise ValueError("all ranks must use the same partial capacity")
                if g == self.rank:
                    self.ptrs.append(self.partial.buffer_ptr)
                    continue
                p = ctypes.c_void_p(0)
                _lib.check(L.mb_peer_open((ctypes.c_ubyte * 64).from_buffer_copy(h), ctypes.byref(p)))
                self.ptrs.append(p)
                self._opened.append(p)

    def _staging(self):
        """world - 1 local slots the peers' partials are pulled into (allocated on first use)."""
        if self._slots is None:
            n = self.partial.nbytes
            self._stage = torch.empty(max(self.world - 1, 1) * n, dtype=torch.uint8, device=self.partial.device)
            base, k = self._stage.data_ptr(), 0
            self._slots = []
            for g in range(self.world):
                if g == self.rank:
                    self._slots.append(self.partial.buffer_ptr)          # my own partial is applied where it is
                else:
                    self._slots.append(ctypes.c_void_p(base + k * n))
                    k += 1
        return self._slots

    def combine(self, layer, pull=True):
        """Applies every rank's partial to layer.data in rank (= time) order.  pull=True (default): ONE kernel first
        copies all peers' rows into local staging slots, reading from every peer at once, then the partials are
        applied from local memory; pull=False: partial g is applied straight out of rank g's memory (every rank reads
        the same owner at the same time: that owner's link is the bottleneck of its step)."""
        dev = self.partial.device
        _rank_barrier(self.group, dev)                      # every rank's fold is complete
        if pull and self.world > 1:
            slots = self._staging()
            arr = ctypes.c_void_p * self.world
            _lib.check(_lib.lib().mb_partial_pull(
                _lib.stream_ptr(dev), arr(*[p.value for p in self.ptrs]), arr(*[p.value for p in slots]), self.world,
                self.rank, self.capacity, self.partial.features))
            for g in range(self.world):                     # rank order = time order
                apply_partial_buffer(layer, slots[g], self.capacity)
        else:
            for g in range(self.world):
                apply_partial_buffer(layer, self.ptrs[g], self.capacity)
        _rank_barrier(self.group, dev)                      # nobody still reads my rows
        self.partial.clear()
        return layer

    def close(self):
        for p in self._opened:
            _lib.lib().mb_peer_close(p)
        self._opened = []


# ---- portable exchange (any backend) -----------------------------------------------------------------------------
def fold_frames(layer, observations, partial=None):
    """Fold `observations` (a contiguous run of frames) into a sparse partial without touching layer.data; returns
    copies of its rows (index, a, b) and leaves the partial empty again."""
    own = partial is None
    if own:
        frames = _num_frames(observations)
        partial = SparsePartial(layer, min(layer.data.shape[0] * layer.data.shape[1] * layer.data.shape[2],
                                           8 * frames * layer.camera_height * layer.camera_width))
    partial.fold(layer, observations)
    layer.check()
    idx, a, b = (t.clone() for t in partial.rows())
    partial.clear()
    return idx, a, b


def apply_partial(layer, idx, a, b):
    """layer.data[idx] = a * layer.data[idx] + b, in place (one rank's partial as tensors)."""
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
    Works on any backend (NCCL on GPUs; gloo in the CPU tests): sizes first, then the three payloads, each padded to
    the largest rank (int64 indices travel as int64)."""
    world = dist.get_world_size(group)
    feat = int(b.shape[1]) if b.dim() == 2 else 0
    n = torch.tensor([idx.numel()], dtype=torch.int64, device=idx.device)
    sizes = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(sizes, n, group=group)
    sizes = [int(x) for x in torch.cat(sizes).tolist()]
    nmax = max(sizes)

    def gather(t, shape, dtype):
        pad = torch.zeros(shape, dtype=dtype, device=idx.device)
        if idx.numel():
            pad[:idx.numel()] = t
        out = [torch.empty_like(pad) for _ in range(world)]
        dist.all_gather(out, pad, group=group)
        return out

    gi = gather(idx.to(torch.int64), (nmax,), torch.int64)
    ga = gather(a.to(torch.float32), (nmax,), torch.float32)
    gb = gather(b.to(torch.float32), (nmax, feat), torch.float32)
    return [(gi[g][:m], ga[g][:m].contiguous(), gb[g][:m].contiguous()) for g, m in enumerate(sizes)]


def update_batch_sharded(layer, local_observations, group=None, partial=None, apply_fn=None, exchange=None):
    """Collective call: every rank passes ITS contiguous chunk of the scene's frames (rank order = time order;
    a rank may pass None or an empty chunk).  On return every rank's layer.data holds the map after all frames,
    equal to sequential fusion up to fp32 re-association (occupancy identical).  `exchange`: a PeerExchange (rows
    are read from the peers' memory inside the apply kernel) or None (all_gather of the rows)."""
    have = local_observations is not None and _num_frames(local_observations) > 0
    if exchange is not None:
        if have:
            exchange.partial.fold(layer, local_observations)
        return exchange.combine(layer)
    if have:
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
