"""Drop-in for /root/reference/mass/nn/base_projection_layer.py (BaseProjectionLayer).

Same constructor kwargs, buffers (`rays`, `data`, `bins_x`, `bins_y`, `bins_z`) and
methods; `update` runs the fused sm_100a kernels of libmassb200 IN PLACE on
`self.data` instead of ~110 eager ATen launches.  The layer must live on a CUDA
device when `update` is called (there is no CPU path).  The coordinate helpers
and `top_down` are not on the hot path and stay plain torch.

Additions over the reference (same semantics, more frames per call):
  * `update_batch(observations)` fuses T frames, in order, in one library call;
  * `exact` (default True) picks the arithmetic of the voxel reduce: the
    reference's operation order (map bitwise equal to the reference CPU path) or
    the per-voxel affine form (<= 1e-5 relative).
"""
from typing import Any, Dict

import numpy as np
import torch
from torch import nn

from mass_b200 import _lib
from mass_b200.nn.projection_layer import ProjectionLayer
from mass_b200.utils.projection import camera_pose, project_camera_rays


def _on_device(x):
    return torch.is_tensor(x) and x.is_cuda


def _as_host_tensor(x):
    return x if torch.is_tensor(x) else torch.from_numpy(np.ascontiguousarray(x))


def _edges(origin, cells, resolution):
    # base_projection_layer.py:164-181: the edge table is whatever ATen's CPU
    # arange produces for these float64 bounds; the kernels only ever read it.
    half = (cells + 1) * resolution / 2
    return torch.arange(origin - half, origin + half - 1e-6, resolution, dtype=torch.float32)


class BaseProjectionLayer(nn.Module, ProjectionLayer):
    """Voxel feature map fed by posed depth + feature images.
    Reference: mass/nn/base_projection_layer.py:15-578."""

    def __init__(self, camera_height: int = 224, camera_width: int = 224,
                 vertical_fov: float = 90.0, map_height: int = 256,
                 map_width: int = 256, map_depth: int = 64,
                 feature_size: int = 1, dtype: torch.dtype = torch.float32,
                 origin_y: float = 0.0, origin_x: float = 0.0,
                 origin_z: float = 0.0, grid_resolution: float = 0.05,
                 interpolation_weight: float = 0.5,
                 initial_feature_map: torch.Tensor = None, exact: bool = True):
        super().__init__()
        if dtype != torch.float32:
            raise ValueError("mass_b200 maps are float32 (the reference's default); got %s" % dtype)
        self.interpolation_weight = interpolation_weight
        self.camera_height, self.camera_width = camera_height, camera_width
        self.vertical_fov = vertical_fov
        self.map_height, self.map_width, self.map_depth = map_height, map_width, map_depth
        self.feature_size = feature_size
        self.origin_x, self.origin_y, self.origin_z = origin_x, origin_y, origin_z
        self.grid_resolution = grid_resolution
        self.exact = exact
        self.min_ray_depth, self.max_ray_depth = 0.0, 10.0     # projection.py:117-118 defaults

        focal = camera_height / 2.0 / np.tan(np.radians(vertical_fov) / 2.0)
        self.register_buffer('rays', project_camera_rays(camera_height, camera_width, focal, focal))
        self.register_buffer('data', torch.zeros(map_height, map_width, map_depth, feature_size, dtype=dtype)
                             if initial_feature_map is None else initial_feature_map)
        self.register_buffer('bins_x', _edges(origin_x, map_width, grid_resolution))
        self.register_buffer('bins_y', _edges(origin_y, map_height, grid_resolution))
        self.register_buffer('bins_z', _edges(origin_z, map_depth, grid_resolution))
        self._ws = None                  # None: the device's shared scratch buffer (_lib.shared_workspace)
        self._updates = 0                # bumped by every kernel that writes the map (see map_state)
        self.frame_graphs = True         # single-frame update(): replay a captured CUDA graph (launch-bound otherwise)
        self._frame_graphs = {}
        self.workspace_limit = None      # optional cap (bytes) on the device scratch buffer of update()

    # -- state ---------------------------------------------------------------------------------
    _TRANSIENT = ("_frame_graphs", "_edge_staging", "_last_ws", "_last_ws_use", "_find_cache", "_match_cache")

    def __getstate__(self):
        """copy.deepcopy / pickle: captured CUDA graphs, events, pinned staging and memoised query results belong to
        this object and its device; a copy starts without them (they are rebuilt on demand)."""
        state = self.__dict__.copy()
        for key in self._TRANSIENT:
            state.pop(key, None)
        state["_frame_graphs"] = {}
        return state

    def reset(self, origin_y: float = 0.0, origin_x: float = 0.0, origin_z: float = 0.0):
        """Zero the map and re-centre it.  Reference: base_projection_layer.py:183-235."""
        self.origin_x, self.origin_y, self.origin_z = origin_x, origin_y, origin_z
        self.data.zero_()
        edges = (_edges(origin_x, self.map_width, self.grid_resolution),
                 _edges(origin_y, self.map_height, self.grid_resolution),
                 _edges(origin_z, self.map_depth, self.grid_resolution))
        bins = (self.bins_x, self.bins_y, self.bins_z)
        if not self.bins_x.is_cuda:
            for b, e in zip(bins, edges):
                b.copy_(e)
            return
        # On the GPU the three tables go through pinned staging and asynchronous copies: a copy from pageable memory
        # blocks the host until everything queued before it (the memset of a map of up to 13.5 GiB just above, the
        # previous episode's kernels) has finished, five layers in a row at every episode start.
        device = self.bins_x.device
        st = getattr(self, "_edge_staging", None)
        if st is None or st["device"] != device:
            st = dict(device=device, pinned=[torch.empty(b.numel(), dtype=torch.float32).pin_memory() for b in bins],
                      done=torch.cuda.Event())
            self._edge_staging = st
        st["done"].synchronize()                   # the previous reset's copies have left the staging buffers
        for b, pinned, e in zip(bins, st["pinned"], edges):
            pinned.copy_(e)
            b.copy_(pinned, non_blocking=True)
        st["done"].record(torch.cuda.current_stream(device))

    def get_feature_map(self):
        return self.data

    def forward(self, observation: Dict[str, Any]):
        self.update(observation)
        return self.get_feature_map()

    # -- the hot path ----------------------------------------------------------------------------
    def _prepare(self, pose, depth, features, class_ids, T):
        """Host side of an update: moves / casts the inputs to device tensors of the library's layout and
        sizes the scratch buffer.  pose [T,12] (CPU or device) f32; depth [T,H,W]; features [T,fh,fw,F] or
        class_ids [T,H,W] int64.  Returns the argument record `_launch` takes."""
        device = _lib.require_cuda(self.data.device)
        if self.data.dtype != torch.float32 or not self.data.is_contiguous():
            raise ValueError("layer.data must be a contiguous float32 tensor")
        H, W, F = self.camera_height, self.camera_width, self.feature_size
        f32 = dict(dtype=torch.float32, device=device)
        depth = torch.as_tensor(depth, **f32).reshape(T, H, W).contiguous()
        pose = pose.reshape(T, 12).to(device, non_blocking=True)
        fh = fw = 0
        if class_ids is not None:
            class_ids = torch.as_tensor(class_ids, dtype=torch.int64, device=device).reshape(T, H, W).contiguous()
        else:
            features = torch.as_tensor(features, **f32)
            if features.dim() == 3:
                features = features[None]
            if features.shape[0] != T or features.shape[-1] != F:
                raise ValueError("features must be [%d, h, w, %d], got %s" % (T, F, tuple(features.shape)))
            fh, fw = features.shape[1], features.shape[2]
            if fh <= 0 or fw <= 0 or H % fh or W % fw:
                # the reference's repeat_interleave would produce a mis-shaped image here
                raise ValueError("feature image %dx%d does not divide the camera %dx%d" % (fh, fw, H, W))
            features = features.contiguous()
        return dict(pose=pose, depth=depth, features=features, class_ids=class_ids, T=T, fh=fh, fw=fw, device=device)

    def map_state(self):
        """Changes whenever the map may have changed: by this package's kernels (which write through the raw
        pointer and so do not bump torch's version counter), by torch in-place ops on `data`, or by rebinding
        `data`.  Derived results (find()'s instance lists) are cached against it."""
        return (self._updates, self.data._version, self.data.data_ptr())

    def mark_dirty(self):
        """Tell the layer its map changed behind its back: a write through `data.data_ptr()` (the free function
        `update_feature_map`, `mb_affine_apply_rows`, any other raw-pointer kernel) or a replay of a CUDA graph the
        CALLER captured around `update_prepared` -- a replay runs no Python, so only the capture bumped the counter.
        Invalidates everything memoised against `map_state()` (find(), predict_scene_differences)."""
        self._updates += 1
        return self

    def _launch(self, prep, fold=None):
        """Device side of an update: enqueues the kernels on the current stream.  No host work besides the
        launches, no allocation once the scratch buffer exists: a call with the same `prep` can be captured
        in a CUDA graph and replayed."""
        device, T, fh, fw = prep["device"], prep["T"], prep["fh"], prep["fw"]
        H, W, F = self.camera_height, self.camera_width, self.feature_size
        if fold is None:
            self._updates += 1
        L = _lib.lib()
        nx, ny, nz = self.bins_x.numel(), self.bins_y.numel(), self.bins_z.numel()
        mode = _lib.MODE_EXACT if (self.exact and fold is None) else _lib.MODE_FAST
        want = L.mb_layer_update_workspace_bytes(H, W, nx, ny, nz, T, F, mode)
        if self.workspace_limit is not None:
            # a smaller scratch buffer makes the library split the call (fewer frames per chunk, more
            # rounds of the feature pass); below the one-frame minimum the call fails
            want = min(want, int(self.workspace_limit))
        ws = (self._ws or _lib.shared_workspace(device)).get(want, device)
        self._last_ws, self._last_ws_use = ws, _lib.note_workspace_use(ws)
        if fold is not None and hasattr(fold, "slot_table"):
            # sparse partial (mass_b200/nn/sharded.py: SparsePartial): rows found / created per touched voxel
            _lib.check(L.mb_layer_fold_sparse(
                _lib.stream_ptr(device), _lib.ptr(self.rays), _lib.ptr(prep["depth"]), _lib.ptr(prep["features"]),
                _lib.ptr(prep["class_ids"]), _lib.ptr(prep["pose"]), T, H, W, fh, fw, F, _lib.ptr(self.bins_x), nx,
                _lib.ptr(self.bins_y), ny, _lib.ptr(self.bins_z), nz, _lib.ptr(fold.slot_table), fold.buffer_ptr,
                int(fold.capacity), float(self.interpolation_weight), float(self.min_ray_depth),
                float(self.max_ray_depth), _lib.ptr(ws), want))
            return self
        if fold is not None:
            partial_b, partial_a = fold
            _lib.check(L.mb_layer_fold(
                _lib.stream_ptr(device), _lib.ptr(self.rays), _lib.ptr(prep["depth"]), _lib.ptr(prep["features"]),
                _lib.ptr(prep["class_ids"]), _lib.ptr(prep["pose"]), T, H, W, fh, fw, F, _lib.ptr(self.bins_x), nx,
                _lib.ptr(self.bins_y), ny, _lib.ptr(self.bins_z), nz, _lib.ptr(partial_b), _lib.ptr(partial_a),
                float(self.interpolation_weight), float(self.min_ray_depth), float(self.max_ray_depth),
                _lib.ptr(ws), want))
            return self
        _lib.check(L.mb_layer_update(
            _lib.stream_ptr(device), _lib.ptr(self.rays), _lib.ptr(prep["depth"]), _lib.ptr(prep["features"]),
            _lib.ptr(prep["class_ids"]), _lib.ptr(prep["pose"]), T, H, W, fh, fw, F, _lib.ptr(self.bins_x), nx,
            _lib.ptr(self.bins_y), ny, _lib.ptr(self.bins_z), nz, _lib.ptr(self.data),
            float(self.interpolation_weight), float(self.min_ray_depth), float(self.max_ray_depth),
            mode, _lib.ptr(ws), want))
        return self

    def _fuse(self, pose, depth, features, class_ids, T, fold=None):
        """fold = (partial_b, partial_a): fold the frames into a partial map instead of self.data
        (frame-sharded scenes, mass_b200/nn/sharded.py)."""
        return self._launch(self._prepare(pose, depth, features, class_ids, T), fold=fold)

    def prepare_batch(self, observations):
        """The host half of update_batch: returns a record of device tensors for `update_prepared`.  Lets a
        caller that replays the same frames (benchmarks, CUDA graphs) pay the host work once."""
        if isinstance(observations, (list, tuple)):
            keys = observations[0].keys()
            observations = {k: torch.stack([torch.as_tensor(o[k]) for o in observations]) for k in keys}
        T = int(torch.as_tensor(observations["yaw"]).reshape(-1).shape[0])
        pose = camera_pose(torch.as_tensor(observations["position"]).reshape(T, 3),
                           torch.as_tensor(observations["yaw"]).reshape(T),
                           torch.as_tensor(observations["elevation"]).reshape(T))
        if "features" in observations:
            return self._prepare(pose, observations["depth"], observations["features"], None, T)
        return self._prepare(pose, observations["depth"], None, observations["class_ids"], T)

    def update_prepared(self, prep):
        """The device half of update_batch (graph-capturable)."""
        return self._launch(prep)

    def check(self):
        """Synchronises and raises if the batched kernels of this layer's last non-exact update flagged an error:
        bit 0 = more accumulate runs than the planned rounds hold (an internal invariant), bit 1 = a class id
        outside [0, feature_size) in a `class_ids` image that was given as a DEVICE tensor (host images are checked
        on the host before the launch; functional.one_hot raises on the same input in the reference).  The bits
        live in the scratch buffer: if another layer has used the device's shared buffer since, they are gone and
        only the synchronisation happens."""
        ws = getattr(self, "_last_ws", None)
        if self.exact or ws is None or getattr(self, "_last_ws_use", None) != _lib.workspace_uses(ws):
            torch.cuda.synchronize(self.data.device)
            return self
        import ctypes
        bits = ctypes.c_uint32(0)
        _lib.check(_lib.lib().mb_layer_update_status(_lib.stream_ptr(self.data.device), _lib.ptr(ws),
                                                     ctypes.byref(bits)))
        if bits.value & 2:
            raise RuntimeError("Class values must be in [0, %d)" % self.feature_size)
        if bits.value & 4:
            raise RuntimeError("libmassb200: the sparse partial is full (capacity too small for the voxels the "
                               "folded frames touch); rows were dropped")
        if bits.value:
            raise RuntimeError("libmassb200: batched update reported error bits 0x%x" % bits.value)
        return self

    def counters(self):
        """Synchronises and returns the device counters of this layer's last batched (non-exact) chunk as a dict:
        items, cells, segments, runs, voxels (distinct voxels touched) and voxel_frames (sum over the chunk's frames
        of the voxels each frame touches).  None if the scratch buffer has been used by another call since."""
        ws = getattr(self, "_last_ws", None)
        if ws is None or getattr(self, "_last_ws_use", None) != _lib.workspace_uses(ws):
            return None
        import ctypes
        c = (ctypes.c_uint32 * 16)()
        _lib.check(_lib.lib().mb_layer_update_counters(_lib.stream_ptr(self.data.device), _lib.ptr(ws), c, 16))
        return dict(items=c[1], cells=c[2], segments=c[3], runs=c[4], error=c[5], voxels=c[6], voxel_frames=c[7])

    def update(self, observation: Dict[str, Any]):
        """Fuse one observation into the map; returns self.
        Reference: base_projection_layer.py:282-343.  Keys: position [3] (x, y, z-up),
        yaw, elevation (radians), depth [H, W, 1], features [h, w, F] (h | H, w | W) --
        or `class_ids` [H, W(, 1)] integer labels standing for one-hot features."""
        pose = camera_pose(observation["position"], observation["yaw"], observation["elevation"])
        features = observation.get("features")
        class_ids = None if features is not None else observation["class_ids"]
        if self.frame_graphs and self.data.is_cuda and not torch.cuda.is_current_stream_capturing():
            # (a caller who is capturing gets the plain launches: they land in the caller's graph)
            return self._update_replayed(pose, observation["depth"], features, class_ids)
        return self._fuse(pose, observation["depth"], features, class_ids, 1)

    def _update_replayed(self, pose, depth, features, class_ids):
        """One frame through a CUDA graph: the ~35 kernels of a single-frame update are launch-bound, so the
        frame is copied into persistent device buffers and the captured launch sequence is replayed."""
        H, W, F = self.camera_height, self.camera_width, self.feature_size
        shape = None if features is None else tuple(torch.as_tensor(features).shape[-3:-1])
        # everything the captured launches bake in as by-value arguments: the reference reads these attributes on
        # every update() (base_projection_layer.py:334-341), so changing one must not replay a stale graph
        key = (shape, self.exact, self.workspace_limit, self.data.data_ptr(),
               torch.cuda.current_stream(self.data.device).cuda_stream, float(self.interpolation_weight),
               float(self.min_ray_depth), float(self.max_ray_depth))
        entry = self._frame_graphs.get(key)
        if entry is None:
            if len(self._frame_graphs) >= 4:
                self._frame_graphs.clear()
            first = self._prepare(pose, depth, features, class_ids, 1)       # validates shapes; persistent copies below
            bufs = {k: (v.clone() if torch.is_tensor(v) else v) for k, v in first.items()}
            self._launch(bufs)                                               # this frame, and sizes the scratch buffer
            torch.cuda.synchronize(self.data.device)
            updates = self._updates
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, capture_error_mode="thread_local"):
                self._launch(bufs)
            self._updates = updates                                          # capturing launches nothing
            self._frame_graphs[key] = (graph, bufs, self._last_ws)           # the scratch buffer must outlive the graph
            return self
        graph, bufs, _ = entry
        f32 = dict(dtype=torch.float32)
        bufs["depth"].copy_(torch.as_tensor(depth, **f32).reshape(1, H, W), non_blocking=True)
        bufs["pose"].copy_(pose.reshape(1, 12), non_blocking=True)
        if class_ids is not None:
            bufs["class_ids"].copy_(torch.as_tensor(class_ids).reshape(1, H, W), non_blocking=True)
        else:
            feats = torch.as_tensor(features, **f32)
            if tuple(feats.shape[-3:]) != tuple(bufs["features"].shape[-3:]):
                raise ValueError("features must be [h, w, %d], got %s" % (F, tuple(feats.shape)))
            bufs["features"].copy_(feats.reshape(bufs["features"].shape), non_blocking=True)
        graph.replay()
        self._updates += 1
        return self

    def update_batch(self, observations, fold=None, check_ids=False):
        """Fuse T observations in order (frames do not commute).  `observations` is a
        list of observation dicts or one dict of stacked arrays with a leading T axis.
        check_ids: frames handed over in HOST memory have their class ids checked by the kernels (the error bits of
        every chunk are collected on the device) and the call raises before it returns."""
        if isinstance(observations, (list, tuple)):
            if len(observations) == 0:
                return self                                        # no frames: the map is unchanged
            keys = observations[0].keys()
            observations = {k: torch.stack([torch.as_tensor(o[k]) for o in observations]) for k in keys}
        T = int(torch.as_tensor(observations["yaw"]).reshape(-1).shape[0])
        if T == 0:
            return self
        pose = camera_pose(torch.as_tensor(observations["position"]).reshape(T, 3),
                           torch.as_tensor(observations["yaw"]).reshape(T),
                           torch.as_tensor(observations["elevation"]).reshape(T))
        features = observations.get("features")
        class_ids = None if features is not None else observations["class_ids"]
        if T > 1 and self.data.is_cuda and not _on_device(observations["depth"]) \
                and not _on_device(features if features is not None else class_ids):
            return self._fuse_from_host(pose, observations["depth"], features, class_ids, T, fold, check_ids=check_ids)
        return self._fuse(pose, observations["depth"], features, class_ids, T, fold=fold)

    host_chunk_bytes = 256 << 20     # staging per chunk when update_batch is handed frames in HOST memory

    def _fuse_from_host(self, pose, depth, features, class_ids, T, fold=None, check_ids=False):
        """T frames that live in host memory (numpy arrays or CPU tensors; pinned memory makes the copies
        asynchronous): the call is cut into chunks of ~host_chunk_bytes, chunk i+1 is copied to one of two device
        staging sets on a copy stream while chunk i is being fused, so the caller sees the PCIe rate of its inputs
        instead of copy + fusion back to back.  Frames still enter the map strictly in order (batches compose)."""
        device = _lib.require_cuda(self.data.device)
        H, W = self.camera_height, self.camera_width
        depth = _as_host_tensor(depth).reshape(T, H, W)
        if depth.dtype != torch.float32:
            depth = depth.to(torch.float32)
        if features is not None:
            feats = _as_host_tensor(features)
            if feats.dim() == 3:
                feats = feats[None]
            if feats.dtype != torch.float32:
                feats = feats.to(torch.float32)
            if feats.shape[0] != T or feats.shape[-1] != self.feature_size:
                raise ValueError("features must be [%d, h, w, %d], got %s" % (T, self.feature_size, tuple(feats.shape)))
            other, name = feats, "features"
        else:
            ids = _as_host_tensor(class_ids).reshape(T, H, W)
            if ids.dtype not in (torch.int64, torch.int32, torch.int16, torch.uint8, torch.int8):
                ids = ids.to(torch.int64)
            other, name = ids, "class_ids"
        per_frame = depth[0].numel() * depth.element_size() + other[0].numel() * other.element_size()
        # ~host_chunk_bytes per chunk, but at least four chunks when the call is small enough to fit fewer: the
        # first chunk's copy is the only one nothing overlaps
        chunk_bytes = min(self.host_chunk_bytes, max(per_frame * T // 4, 32 << 20))
        chunk = int(max(1, min(T, chunk_bytes // max(per_frame, 1))))
        errors = torch.zeros((), dtype=torch.int32, device=device) if check_ids and name == "class_ids" else None
        st = _lib.host_staging(device)
        main = torch.cuda.current_stream(device)
        for slot in st.slots:
            slot["free"].record(main)                  # staging sets may be overwritten once earlier work is done
        for i, s in enumerate(range(0, T, chunk)):
            e = min(s + chunk, T)
            slot = st.slots[i % 2]
            d_buf = st.buffer(i % 2, "depth", (e - s, H, W), torch.float32)
            o_buf = st.buffer(i % 2, name, (e - s,) + tuple(other.shape[1:]), other.dtype)
            with torch.cuda.stream(st.stream):
                st.stream.wait_event(slot["free"])
                d_buf.copy_(depth[s:e], non_blocking=True)
                o_buf.copy_(other[s:e], non_blocking=True)
                slot["ready"].record(st.stream)
            main.wait_event(slot["ready"])
            if name == "features":
                self._fuse(pose[s:e], d_buf, o_buf, None, e - s, fold=fold)
            else:
                self._fuse(pose[s:e], d_buf, None, o_buf if o_buf.dtype == torch.int64 else o_buf.to(torch.int64),
                           e - s, fold=fold)
                if errors is not None and not self.exact:
                    # the chunk's sticky error word (bit 1: a class id outside [0, F)), OR-ed on the device
                    errors |= self._last_ws[:64].view(torch.int32)[_lib.CNT_ERROR]
            slot["free"].record(main)
        if errors is not None and int(errors.item()) & 2:
            raise RuntimeError("Class values must be in [0, %d)" % self.feature_size)
        return self

    # -- whole-map readers next to the path ------------------------------------------------
    def column_summary(self, depth_slice: slice = None, obstacle_threshold: float = 0.0,
                       want_amax: bool = True, want_blocked: bool = True):
        """One pass over the map: (amax [S0, S1, F] = data.amax(dim=2), blocked [S0, S1] bool =
        (norm(data, p=1, dim=3) > obstacle_threshold)[:, :, depth_slice].any(dim=2)).
        Reference: agent.py:330-331 (policy input), mass/navigation_policy.py:207-216 (obstacles)."""
        data = self.data
        device = _lib.require_cuda(data.device)
        S0, S1, S2, F = data.shape
        z_lo, z_hi, step = (depth_slice or slice(None)).indices(S2)
        if step != 1:
            raise ValueError("depth_slice must have step 1")
        z_hi = max(z_hi, z_lo)
        amax = torch.empty(S0, S1, F, dtype=torch.float32, device=device) if want_amax else None
        blocked = torch.empty(S0, S1, dtype=torch.uint8, device=device) if want_blocked else None
        _lib.check(_lib.lib().mb_column_summary(_lib.stream_ptr(device), _lib.ptr(data), S0, S1, S2, F, z_lo, z_hi,
                                                float(obstacle_threshold), _lib.ptr(amax), _lib.ptr(blocked)))
        return amax, (blocked.bool() if blocked is not None else None)

    # -- rendering + coordinate helpers (not on the hot path; plain torch) --------------------------
    def top_down(self, depth_slice: slice = slice(0, 32)):
        """Features of the top-most non-empty voxel per (y, x) column.
        Reference: base_projection_layer.py:345-379."""
        data = self.data
        S0, S1, S2, F = data.shape
        z_lo, z_hi, step = (depth_slice or slice(None)).indices(S2)
        if not data.is_cuda or step != 1 or z_hi <= z_lo:
            # not the device path (a CPU copy of the layer, a strided or empty slice): the reference's expression
            vol = data if depth_slice is None else data[:, :, depth_slice]
            filled = (vol != 0).any(dim=-1, keepdim=True).to(vol.dtype)
            top = (filled.cumsum(dim=-2) * filled).argmax(dim=-2, keepdim=True)
            return torch.gather(vol, -2, top.expand(*vol.shape[:-2], 1, vol.shape[-1])).squeeze(-2)
        out = torch.empty(S0, S1, F, dtype=torch.float32, device=data.device)
        _lib.check(_lib.lib().mb_top_down(_lib.stream_ptr(data.device), _lib.ptr(data), S0, S1, S2, F, z_lo, z_hi,
                                          _lib.ptr(out)))
        return out

    def _centre_span(self):
        lo = torch.stack([(b[0] + b[1]) / 2 for b in (self.bins_x, self.bins_y, self.bins_z)])
        hi = torch.stack([(b[-1] + b[-2]) / 2 for b in (self.bins_x, self.bins_y, self.bins_z)])
        return lo, hi

    def clamp_to_world(self, coords):
        """Reference: base_projection_layer.py:381-413."""
        coords = torch.as_tensor(coords, dtype=torch.float32, device=self.data.device)
        lo, hi = self._centre_span()
        k = coords.shape[-1]
        lead = [1] * (coords.dim() - 1)
        return coords.clamp(min=lo[:k].view(*lead, k), max=hi[:k].view(*lead, k))

    def clamp_to_map(self, coords):
        """Reference: base_projection_layer.py:415-450."""
        coords = torch.as_tensor(coords, dtype=coords.dtype, device=self.data.device)
        k = coords.shape[-1]
        lead = [1] * (coords.dim() - 1)
        hi = torch.tensor([self.map_width - 1, self.map_height - 1, self.map_depth - 1],
                          dtype=coords.dtype, device=coords.device)[:k].view(*lead, k)
        return coords.clamp(min=torch.zeros_like(hi), max=hi)

    def cell_centres(self):
        """Per-axis world coordinate of every cell centre (y already flipped)."""
        mx = (self.bins_x[:-1] + self.bins_x[1:]) / 2
        my = (self.bins_y[:-1] + self.bins_y[1:]).flip(-1) / 2
        mz = (self.bins_z[:-1] + self.bins_z[1:]) / 2
        return mx, my, mz

    def _transform(self, coords, to_world):
        """Batched coordinate transform on the device (mb_map_to_world / mb_world_to_map): any leading shape, last
        dimension 2 (xy) or 3 (xyz)."""
        device = self.data.device
        coords = torch.as_tensor(coords)
        if to_world:
            # clamp_to_map runs in the caller's dtype, then the reference casts to float32 (lines 471-472)
            coords = self.clamp_to_map(coords.to(device)).to(torch.float32)
        coords = coords.to(device=device, dtype=torch.float32).contiguous()
        k = int(coords.shape[-1])
        if k not in (2, 3):
            raise ValueError("coordinates must end in a dimension of 2 (xy) or 3 (xyz), got %s" % (tuple(coords.shape),))
        n = coords.numel() // k
        out = torch.empty(coords.shape, dtype=torch.float32 if to_world else torch.int64, device=device)
        fn = _lib.lib().mb_map_to_world if to_world else _lib.lib().mb_world_to_map
        _lib.check(fn(_lib.stream_ptr(device), _lib.ptr(coords), n, k, _lib.ptr(self.bins_x), self.bins_x.numel(),
                      _lib.ptr(self.bins_y), self.bins_y.numel(), _lib.ptr(self.bins_z), self.bins_z.numel(),
                      _lib.ptr(out)))
        return out

    def map_to_world(self, coords):
        """Map (x, y, z) cell coordinates (fractional allowed) -> world.
        Reference: base_projection_layer.py:452-511.  One kernel for the whole batch when the layer is on the GPU."""
        if self.data.is_cuda:
            return self._transform(coords, to_world=True)
        coords = self.clamp_to_map(coords).to(torch.float32)
        base = coords.floor()
        cell = base.to(torch.int64)
        frac = coords - base
        out = []
        for axis, (mid, size) in enumerate(zip(self.cell_centres(),
                                               (self.map_width, self.map_height, self.map_depth))):
            if axis >= coords.shape[-1]:
                break
            left = mid[cell[..., axis]]
            right = mid[(cell[..., axis] + 1).clamp(0, size - 1)]
            out.append(left + (right - left) * frac[..., axis])
        return torch.stack(out, dim=-1)

    def world_to_map(self, coords):
        """World -> integer map (x, y, z) cells.  Reference: base_projection_layer.py:513-547.  One kernel for the
        whole batch when the layer is on the GPU."""
        if self.data.is_cuda:
            return self._transform(coords, to_world=False)
        coords = self.clamp_to_world(coords)
        cells = [torch.bucketize(coords[..., 0].contiguous(), self.bins_x, right=True) - 1,
                 self.bins_y.numel() - torch.bucketize(coords[..., 1].contiguous(), self.bins_y, right=True) - 1]
        if coords.shape[-1] == 3:
            cells.append(torch.bucketize(coords[..., 2].contiguous(), self.bins_z, right=True) - 1)
        return torch.stack(cells, dim=-1)

    def visualize(self, obs: Dict[str, Any], depth_slice: slice = slice(0, 32)):
        """White = empty column, black = occupied.  Reference: base_projection_layer.py:549-578."""
        vol = self.data if depth_slice is None else self.data[:, :, depth_slice]
        occ = (vol != 0).any(dim=-1, keepdim=True).to(torch.float32).detach().cpu().numpy()
        return 1.0 - np.tile(occ, (1, 1, 3))
