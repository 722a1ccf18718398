"""ctypes binding of libmassb200.so (the C ABI declared in include/massb200.h).

There is no CPU fallback and no alternative backend: if the shared library is
missing it is built in-tree with nvcc for sm_100a, and if that is impossible the
import of any hot-path function fails loudly.
"""
import ctypes
import fcntl
import hashlib
import os
import subprocess

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
SO_PATH = os.path.join(CSRC, "libmassb200.so")
STAMP_PATH = SO_PATH + ".srchash"

MODE_EXACT = 0
MODE_FAST = 1
CNT_ERROR = 5        # index of the sticky error word among the batched pipeline's device counters (kernels.cuh)

_lib = None

_vp = ctypes.c_void_p
_i64 = ctypes.c_int64
_i32 = ctypes.c_int
_f32 = ctypes.c_float
_sz = ctypes.c_size_t


def _sources():
    srcs = sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh")) or f == "Makefile")
    srcs.append(os.path.join(os.path.dirname(_HERE), "include", "massb200.h"))
    return srcs


def source_hash():
    """sha256 over every source the library is built from.  Compared with the stamp written next to the .so, so a
    stale binary is never loaded (file times do not survive a snapshot copy; contents do)."""
    h = hashlib.sha256()
    for path in _sources():
        h.update(os.path.basename(path).encode())
        with open(path, "rb") as f:
            h.update(f.read())
    return h.hexdigest()


def is_stale():
    if not os.path.exists(SO_PATH) or not os.path.exists(STAMP_PATH):
        return True
    with open(STAMP_PATH) as f:
        return f.read().strip() != source_hash()


def build(force=False):
    """Compile every CUDA source for sm_100a (see csrc/Makefile); a no-op when the library matches the sources.
    Concurrent callers (one process per GPU) serialise on a lock file."""
    if not force and not is_stale():
        return SO_PATH
    with open(os.path.join(CSRC, ".build.lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if force or is_stale():
                proc = subprocess.run(["make", "-C", CSRC, "-j8", "-B"], stdout=subprocess.PIPE,
                                      stderr=subprocess.STDOUT, text=True)
                if proc.returncode != 0:
                    raise RuntimeError("building libmassb200.so failed:\n" + proc.stdout[-4000:])
                with open(STAMP_PATH, "w") as f:
                    f.write(source_hash())
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)
    return SO_PATH


_SIGNATURES = {
    "mb_last_error": (ctypes.c_char_p, []),
    "mb_version": (_i32, []),
    "mb_launch_count": (ctypes.c_uint64, []),
    "mb_transform_rays": (_i32, [_vp, _vp, _i64, _vp, _vp]),
    "mb_bin_rays_workspace_bytes": (_sz, [_i64]),
    "mb_bin_rays": (_i32, [_vp, _vp, _i32, _vp, _i32, _vp, _i32, _vp, _vp, _vp, _i64, _f32, _f32,
                           _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _sz]),
    "mb_update_feature_map_workspace_bytes": (_sz, [_i64, _i32, _i32, _i32]),
    "mb_update_feature_map": (_i32, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _i32, _vp, _i32, _i32,
                                     _i32, _f32, _i32, _vp, _sz]),
    "mb_layer_update_workspace_bytes": (_sz, [_i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32]),
    "mb_layer_update_min_workspace_bytes": (_sz, [_i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32]),
    "mb_layer_fold": (_i32, [_vp, _vp, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _vp, _i32,
                             _vp, _i32, _vp, _i32, _vp, _vp, _f32, _f32, _f32, _vp, _sz]),
    "mb_affine_apply_rows": (_i32, [_vp, _vp, _i32, _vp, _vp, _vp, _i64]),
    "mb_partial_buffer_bytes": (_sz, [ctypes.c_uint32, _i32]),
    "mb_partial_buffer_layout": (_i32, [ctypes.c_uint32, _i32, _vp]),
    "mb_partial_reset": (_i32, [_vp, _vp, _i64, _vp]),
    "mb_partial_clear": (_i32, [_vp, _vp, _vp, ctypes.c_uint32, _i32]),
    "mb_layer_fold_sparse": (_i32, [_vp, _vp, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _vp, _i32,
                                    _vp, _i32, _vp, _i32, _vp, _vp, ctypes.c_uint32, _f32, _f32, _f32, _vp, _sz]),
    "mb_affine_apply_partial": (_i32, [_vp, _vp, _i32, _vp, ctypes.c_uint32]),
    "mb_partial_pull": (_i32, [_vp, _vp, _vp, _i32, _i32, ctypes.c_uint32, _i32]),
    "mb_peer_alloc": (_i32, [_sz, _vp]),
    "mb_peer_free": (_i32, [_vp]),
    "mb_peer_export": (_i32, [_vp, _vp]),
    "mb_peer_open": (_i32, [_vp, _vp]),
    "mb_peer_close": (_i32, [_vp]),
    "mb_profile_stages": (_i32, [_i32]),
    "mb_profile_read": (_i32, [_vp, _i32]),
    "mb_layer_update_status": (_i32, [_vp, _vp, _vp]),
    "mb_layer_update_counters": (_i32, [_vp, _vp, _vp, _i32]),
    "mb_class_presence_workspace_bytes": (_sz, [_i32, _i32, _i32, _i32]),
    "mb_class_presence": (_i32, [_vp, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _f32, _vp, _vp, _sz]),
    "mb_instance_pool": (_i32, [_vp, _vp, _i32, _vp, _i32, _i32, _i32, _i32, _i32, _vp, _i32, _vp, _vp, _vp, _vp]),
    "mb_column_summary": (_i32, [_vp, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _f32, _vp, _vp]),
    "mb_masks_to_ids": (_i32, [_vp, _vp, _vp, _vp, _i32, _i64, _i32, _f32, _vp]),
    "mb_top_down": (_i32, [_vp, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _vp]),
    "mb_world_to_map": (_i32, [_vp, _vp, _i64, _i32, _vp, _i32, _vp, _i32, _vp, _i32, _vp]),
    "mb_map_to_world": (_i32, [_vp, _vp, _i64, _i32, _vp, _i32, _vp, _i32, _vp, _i32, _vp]),
    "mb_navigable_area": (_i32, [_vp, _vp, _i32, _i32, _i32, _vp]),
    "mb_nav_graph_lattice": (_i32, [_vp, _vp, _i32, _i32, _i32, _i32, _i32, _vp, _vp]),
    "mb_nav_rects_clear": (_i32, [_vp, _vp, _i32, _i32, _vp, _i32, _vp]),
    "mb_pairwise_l2": (_i32, [_vp, _vp, _i32, _vp, _i32, _i32, _vp]),
    "mb_cosine_best_match": (_i32, [_vp, _vp, _i32, _vp, _i32, _i32, _vp, _vp]),
    "mb_cosine_best_match_tc_workspace_bytes": (_sz, [_i32, _i32, _i32]),
    "mb_cosine_best_match_tc": (_i32, [_vp, _vp, _i32, _vp, _i32, _i32, _vp, _vp, _vp, _sz]),
    "mb_lsap_workspace_bytes": (_sz, [_i32, _i32]),
    "mb_lsap": (_i32, [_vp, _vp, _vp, _i32, _i32, _vp, _vp, _vp, _vp, _sz]),
    "mb_layer_update": (_i32, [_vp, _vp, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _vp, _i32,
                               _vp, _i32, _vp, _i32, _vp, _f32, _f32, _f32, _i32, _vp, _sz]),
}


def lib():
    """Loads (building first if needed) the shared library; raises if impossible."""
    global _lib
    if _lib is None:
        override = os.environ.get("MASSB200_LIB")         # tuning aid: a variant build of the same sources
        if override:
            L = ctypes.CDLL(override)
        else:
            build()                        # no-op unless a source changed since the library was built
            L = ctypes.CDLL(SO_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(L, name)          # AttributeError here = header/library mismatch
            fn.restype, fn.argtypes = res, args
        _lib = L
    return _lib


def exported_symbols():
    return sorted(_SIGNATURES)


def check(rc):
    if rc != 0:
        msg = lib().mb_last_error().decode("utf-8", "replace")
        if rc == 1:
            raise ValueError("libmassb200: " + msg)
        raise RuntimeError("libmassb200 (status %d): %s" % (rc, msg))


def stream_ptr(device):
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def ptr(t):
    """Device pointer of a tensor that must already be CUDA + contiguous."""
    if t is None:
        return ctypes.c_void_p(0)
    if not t.is_cuda:
        raise ValueError("mass_b200 kernels need CUDA tensors (there is no CPU path), got %s" % t.device)
    if not t.is_contiguous():
        raise ValueError("tensor must be contiguous")
    return ctypes.c_void_p(t.data_ptr())


def require_cuda(device):
    device = torch.device(device)
    if device.type != "cuda":
        raise RuntimeError("mass_b200 runs on CUDA (sm_100a) only; the layer is on %s. "
                           "Move it with .cuda() first -- there is no CPU fallback." % device)
    return device


class Workspace:
    """A grow-only device scratch buffer owned by the caller side (torch allocator)."""

    def __init__(self):
        self.buf = None

    def get(self, nbytes, device):
        if self.buf is None or self.buf.numel() < nbytes or self.buf.device != device:
            self.buf = torch.empty(int(nbytes), dtype=torch.uint8, device=device)
        return self.buf


_ws_uses = {}


def note_workspace_use(buf):
    """Counts the library calls that were handed `buf` as scratch (its counters are overwritten by each)."""
    key = buf.data_ptr()
    _ws_uses[key] = _ws_uses.get(key, 0) + 1
    return _ws_uses[key]


def workspace_uses(buf):
    return _ws_uses.get(buf.data_ptr(), 0)


_shared = {}


def shared_workspace(device):
    """One grow-only scratch buffer per device, shared by every layer on it (an episode keeps five maps; each
    batched update wants GBs of scratch).  Sharing is safe while the updates of a device are enqueued on one
    stream, which is how the reference drives its layers; give a layer its own `Workspace()` (layer._ws) to update
    it from another stream."""
    key = torch.device(device)
    if key.type == "cuda" and key.index is None:
        key = torch.device("cuda", torch.cuda.current_device())
    if key not in _shared:
        _shared[key] = Workspace()
    return _shared[key]



class HostStaging:
    """Two sets of device staging buffers, a copy stream and the events that order them: what update_batch needs to
    overlap the host->device copy of chunk i+1 with the fusion of chunk i when it is handed HOST frames."""

    def __init__(self, device):
        self.device = device
        self.stream = torch.cuda.Stream(device=device)
        self.slots = [dict(bufs={}, ready=torch.cuda.Event(), free=torch.cuda.Event()) for _ in range(2)]

    def buffer(self, slot, name, shape, dtype):
        """Grow-only device tensor `name` of slot `slot`, viewed as `shape`."""
        need = 1
        for d in shape:
            need *= int(d)
        cur = self.slots[slot]["bufs"].get(name)
        if cur is None or cur.dtype != dtype or cur.numel() < need:
            cur = torch.empty(need, dtype=dtype, device=self.device)
            self.slots[slot]["bufs"][name] = cur
        return cur[:need].view(*shape)


_staging = {}


def host_staging(device):
    key = torch.device(device)
    if key.type == "cuda" and key.index is None:
        key = torch.device("cuda", torch.cuda.current_device())
    if key not in _staging:
        _staging[key] = HostStaging(key)
    return _staging[key]
