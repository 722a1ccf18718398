"""bench_configs.py -- the BASELINE configurations bench.py's default line is NOT quoted on.

  c4  large scene: 720x1280 depth + 54-class probabilities, 0.02 m grid (960x960x240 map, 47.8 GB resident per
      GPU), the frames of ONE scene split into contiguous chunks across the ranks, every rank folding its chunk
      ring by ring (32 resident frames) into a sparse partial, partials combined in rank order by the ordered affine
      apply that reads the peers' rows over NVLink (mass_b200/nn/sharded.py).  STRONG scaling: the scene is fixed.
  c3  full episode pair incl. the MATCH stage: walkthrough + unshuffle maps (occupancy, 2 x semantic from class ids,
      2 x 256-d instance features at 56x56), then the agent's predict_scene_differences loop over ~200 objects
      (mass/utils/experimentation.py:235-311, agent.py:424-465).

Both print one JSON line (rank 0) in bench.py's format; `--check` adds a comparison with the CPU oracle on a frame
subset (and, for c4, with sequential fusion on one GPU).  Synthetic inputs are rendered on the GPU; generation is
outside every timed region.
"""
import json
import os
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))

C4 = dict(camera_height=720, camera_width=1280, vertical_fov=90.0, map_height=960, map_width=960, map_depth=240,
          feature_size=54, grid_resolution=0.02, interpolation_weight=0.5)
C4_FRAMES = 4096
C4_RING = 32                    # resident frames per GPU and fold call
C4_PARTIAL_ROWS = 8_000_000     # capacity of a rank's sparse partial (228 B per row)


# ---- synthetic frames on the GPU -------------------------------------------------------------------------------------
class BoxRoomRenderer:
    """Planar depth of the box-room (+ optional object boxes) for a pinhole camera, float64 math on the device:
    the same scene model as mass_b200/utils/synthetic.py, batched."""

    def __init__(self, H, W, dev, boxes=None, classes=None, vertical_fov=90.0):
        from mass_b200.utils import synthetic
        self.synthetic, self.dev, self.H, self.W = synthetic, dev, H, W
        self.rays = torch.tensor(synthetic.camera_rays(H, W, vertical_fov), device=dev)       # [H,W,3] f64
        self.lo = torch.tensor(synthetic.ROOM_LO, device=dev)
        self.hi = torch.tensor(synthetic.ROOM_HI, device=dev)
        self.boxes, self.classes = boxes, classes

    def frame(self, position, yaw, elevation):
        rot = torch.tensor(self.synthetic._rotation(yaw, elevation), device=self.dev)
        r = self.rays @ rot.T
        o = torch.tensor(np.asarray(position, np.float64), device=self.dev)
        far = torch.where(r > 0, self.hi, self.lo)
        t_wall = ((far - o) / r).amin(-1)
        if self.boxes is None:
            return t_wall.to(torch.float32), None
        inv = 1.0 / r[..., None, :]
        t0, t1 = (self.boxes[:, :3] - o) * inv, (self.boxes[:, 3:] - o) * inv
        tn, tf = torch.minimum(t0, t1).amax(-1), torch.maximum(t0, t1).amin(-1)
        tn = torch.where((tn <= tf) & (tn > 0), tn, torch.full_like(tn, float("inf")))
        tb, k = tn.min(-1)
        hit = torch.where(tb < t_wall, k, torch.full_like(k, -1))
        return torch.minimum(t_wall, tb).to(torch.float32), hit


def c4_ring(renderer, t0, n, total, F, dev):
    """Frames t0 .. t0+n-1 of the c4 scene (an orbit of `total` frames): poses, depth [n,H,W,1], probabilities
    [n,H,W,F] = softmax(4 randn) at 1/8 resolution, nearest up-sampled (seed 1000 + t0)."""
    H, W = renderer.H, renderer.W
    pos, yaw, elev, depth = [], [], [], []
    for t in range(t0, t0 + n):
        p, y, e = renderer.synthetic.boxroom_pose(t, total)
        d, _ = renderer.frame(p, y, e)
        pos.append(p), yaw.append(y), elev.append(e), depth.append(d[..., None])
    g = torch.Generator(device=dev).manual_seed(1000 + t0)
    low = torch.softmax(4 * torch.randn(n, H // 8, W // 8, F, device=dev, generator=g), dim=-1)
    probs = low.repeat_interleave(8, 1).repeat_interleave(8, 2).contiguous()
    return dict(position=np.stack(pos), yaw=np.array(yaw, np.float32), elevation=np.array(elev, np.float32),
                depth=torch.stack(depth), features=probs)


def _host_ram_gb():
    try:
        import psutil
        return psutil.virtual_memory().available / 1e9
    except Exception:
        return 0.0


# ---- c4 ----------------------------------------------------------------------------------------------------------------
def c4_sharded(R, frames_total=C4_FRAMES, brief=False, check=False, ring=C4_RING):
    """One pass over the c4 scene on R.world ranks.  Returns the result dict on every rank."""
    from mass_b200.nn import sharded
    from mass_b200.nn.base_projection_layer import BaseProjectionLayer
    from mass_b200.utils import synthetic

    dev, world, rank = R.dev, R.world, R.rank
    H, W, F = C4["camera_height"], C4["camera_width"], C4["feature_size"]
    frames_total = int(frames_total) // (world * ring) * (world * ring)
    if frames_total <= 0:
        raise ValueError("c4 needs at least %d frames on %d ranks" % (world * ring, world))
    per_rank = frames_total // world
    first = rank * per_rank
    layer = BaseProjectionLayer(exact=False, **C4, **synthetic.MAP_ORIGIN).to(dev)
    renderer = BoxRoomRenderer(H, W, dev)
    ex = sharded.PeerExchange(layer, C4_PARTIAL_ROWS) if world > 1 else None
    ev = lambda: torch.cuda.Event(enable_timing=True)                                     # noqa: E731

    def one_pass(timed):
        fold_ms, marks = 0.0, []
        for c in range(per_rank // ring):
            obs = c4_ring(renderer, first + c * ring, ring, frames_total, F, dev)
            prep = layer.prepare_batch(obs)                      # pose upload etc.: host work outside the timed span
            e0, e1 = ev(), ev()
            e0.record()
            if ex is None:
                layer.update_prepared(prep)                      # one GPU: sequential fusion straight into the map
            else:
                layer._launch(prep, fold=ex.partial)             # this rank's next ring onto its sparse partial
            e1.record()
            marks.append((e0, e1))
            del obs, prep
        c0, c1 = ev(), ev()
        R.barrier()                                              # all ranks start the combine together
        c0.record()
        if ex is not None:
            rows = ex.partial.count_view.clone()                 # (device copy; read after the timed span)
            ex.combine(layer, pull=os.environ.get("MASSB200_C4_DIRECT") != "1")
        c1.record()
        torch.cuda.synchronize()
        fold_ms = sum(a.elapsed_time(b) for a, b in marks)
        comb_ms = c0.elapsed_time(c1)
        nrows = int(rows.item()) if ex is not None else 0
        return fold_ms, comb_ms, nrows

    one_pass(False)                                              # warm-up: scratch buffers, NCCL channels, peer mappings
    layer.check()
    layer.data.zero_()
    torch.cuda.synchronize()
    fold_ms, comb_ms, nrows = one_pass(True)
    layer.check()
    fold_max, comb_max = R.max_ms(fold_ms), R.max_ms(comb_ms)
    total_ms = fold_max + comb_max          # the combine starts after a barrier: the slowest fold + the slowest combine
    rows_all = [0] * world
    if world > 1:
        t = torch.zeros(world, dtype=torch.float64, device=dev)
        t[rank] = nrows
        R.dist.all_reduce(t)
        rows_all = [int(x) for x in t.tolist()]
    row_bytes = 8 + 4 + 4 * F
    link_bytes = [sum(rows_all[g] for g in range(world) if g != r) * row_bytes for r in range(world)]
    # every replica must hold the same map: occupancy count and a checksum agree across ranks
    occ = float((layer.data[..., 0] != 0).sum().item())
    chk = float(layer.data[:, :, :, ::9].double().sum().item())
    same = True
    if world > 1:
        t = torch.tensor([occ, chk], dtype=torch.float64, device=dev)
        lo, hi = t.clone(), t.clone()
        R.dist.all_reduce(lo, op=R.dist.ReduceOp.MIN)
        R.dist.all_reduce(hi, op=R.dist.ReduceOp.MAX)
        same = bool(torch.equal(lo, hi))
    out = {"workload": "c4: %d frames of 720x1280 depth + 54-class probabilities of ONE scene into a 960x960x240 map at "
                       "0.02 m; contiguous chunks of %d frames per rank, %d resident frames per fold call"
                       % (frames_total, per_rank, ring),
           "value": frames_total / (total_ms * 1e-3), "unit": "frames/s", "n_gpus": world, "scaling": "strong",
           "frames": frames_total, "ms": {"total_max_over_ranks": total_ms, "fold": fold_max, "combine": comb_max},
           "combine_share": comb_max / total_ms if total_ms > 0 else 0.0,
           "exchange": (("peer memory over NVLink: each rank's sparse partial is read out of its owner's HBM by the apply "
                         "kernel of every other rank (mb_affine_apply_partial)" if os.environ.get("MASSB200_C4_DIRECT") == "1" else
                         "peer memory over NVLink: one kernel per rank pulls the sparse partials of all peers at once "
                         "(mb_partial_pull, row counts read on the device), then they are applied in rank order")
                        + ", two stream-ordered NCCL barriers around it")
           if world > 1 else "none (one GPU: sequential fusion into the map)",
           "partial_rows_per_rank": rows_all, "nvlink_bytes_read_per_rank": link_bytes,
           "nvlink_gbs_per_rank": (max(link_bytes) / (comb_max * 1e-3) / 1e9) if (world > 1 and comb_max > 0) else None,
           "replicas_identical": same, "occupied_voxels": int(occ)}
    if check:
        out["check"] = c4_check(R, layer, renderer, frames_total, ring)
    if ex is not None:
        ex.close()
    del layer, ex
    torch.cuda.empty_cache()
    return out


def c4_check(R, layer, renderer, frames_total, ring):
    """(1) rank 0 fuses ALL frames sequentially into a second map and compares it with its sharded replica
    (occupancy identical, values within 2e-5: both are within 1e-5 of the reference's order of operations);
    (2) the CPU oracle fuses the first two frames of the scene at full map size and is compared with the same two
    frames on the GPU (occupancy bit-exact, values within 1e-5) -- skipped, and said so, if the host lacks the RAM."""
    from mass_b200.nn.base_projection_layer import BaseProjectionLayer
    from mass_b200.utils import synthetic
    dev, F = R.dev, C4["feature_size"]
    res = {}
    if R.rank == 0:
        ref = BaseProjectionLayer(exact=False, **C4, **synthetic.MAP_ORIGIN).to(dev)
        for c in range(frames_total // ring):
            ref.update_batch(c4_ring(renderer, c * ring, ring, frames_total, F, dev))
        torch.cuda.synchronize()
        occ_equal, rel_ok, worst = True, True, 0.0
        for y in range(0, ref.data.shape[0], 32):                # slab by slab: no map-sized temporaries
            r, l = ref.data[y:y + 32], layer.data[y:y + 32]
            occ_equal = occ_equal and bool(torch.equal((r != 0).any(-1), (l != 0).any(-1)))
            d = (r - l).abs()
            rel_ok = rel_ok and bool((d <= 2e-5 * r.abs()).all())
            nz = r != 0
            if bool(nz.any()):
                worst = max(worst, float((d[nz] / r.abs()[nz]).max()))
        res["sharded_vs_sequential_gpu"] = {"occupancy_identical": occ_equal, "values_within_2e-5": rel_ok,
                                            "worst_rel": worst}
        del ref, r, l, d, nz                                     # (the slab views keep the whole map alive)
        torch.cuda.empty_cache()
        need = 4.0 * 960 * 960 * 240 * 54 / 1e9 * 1.3
        if _host_ram_gb() > need + 20:
            from oracle import oracle
            obs = c4_ring(renderer, 0, 2, frames_total, F, dev)
            gpu = BaseProjectionLayer(exact=False, **C4, **synthetic.MAP_ORIGIN).to(dev).update_batch(obs)
            orc = oracle.OracleLayer(nthreads=os.cpu_count() or 8, **C4, **synthetic.MAP_ORIGIN)
            for t in range(2):
                orc.update(dict(position=obs["position"][t], yaw=obs["yaw"][t], elevation=obs["elevation"][t],
                                depth=obs["depth"][t].cpu().numpy(), features=obs["features"][t].cpu().numpy()))
            occ_equal, rel_ok = True, True
            for y in range(0, 960, 24):
                o = torch.from_numpy(orc.data[y:y + 24]).to(dev)
                g = gpu.data[y:y + 24]
                occ_equal = occ_equal and bool(torch.equal((o != 0).any(-1), (g != 0).any(-1)))
                rel_ok = rel_ok and bool(((o.double() - g.double()).abs() <= 1e-5 * o.double().abs()).all())
            res["gpu_vs_oracle_2_frames_full_map"] = {"occupancy_bit_exact": occ_equal, "values_within_1e-5": rel_ok}
            del gpu, orc, o, g
            torch.cuda.empty_cache()
        else:
            res["gpu_vs_oracle_2_frames_full_map"] = "skipped: host has %.0f GB free, the oracle's map needs %.0f GB" % (
                _host_ram_gb(), need)
    R.barrier()
    return res


def run_c4(args, R):
    frames = args.c4_frames if args.c4_frames > 0 else C4_FRAMES
    out = c4_sharded(R, frames_total=frames, check=args.check)
    if R.rank == 0:
        line = {"metric": "RGB-D frames/sec fused into semantic voxel map", "value": out["value"], "unit": "frames/s",
                "n_gpus": R.world, "steps": 1, "warmup": 1, "ms_per_step": out["ms"]["total_max_over_ranks"],
                "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": out["workload"]}, "c4": out}
        print(json.dumps(line))


# ---- c3 ----------------------------------------------------------------------------------------------------------------
C3_MAP = dict(vertical_fov=90.0, map_height=384, map_width=384, map_depth=96, grid_resolution=0.05,
              interpolation_weight=0.5)
C3_OBJECTS = 200
C3_FRAMES = (250, 500)          # walkthrough, unshuffle
C3_MATCH_KW = dict(confidence_threshold=0.0, contour_padding=0, contour_threshold=0.0, distance_threshold=0.05)


def c3_boxes(shifted, dev):
    """~200 axis-aligned object boxes of random class 1..53 in the box-room; a subset is shifted for the unshuffle
    pass (SURVEY.md 8d, config 3 proposal)."""
    rng = np.random.default_rng(23)
    c = rng.uniform([-3.7, -2.7, 0.0], [3.7, 2.7, 1.6], (C3_OBJECTS, 3))
    s = rng.uniform(0.12, 0.3, (C3_OBJECTS, 3))
    cls = rng.integers(1, 54, C3_OBJECTS)
    if shifted:
        moved = rng.choice(C3_OBJECTS, 25, replace=False)
        c[moved, :2] += rng.uniform(0.4, 0.8, (25, 2)) * rng.choice([-1, 1], (25, 2))
    lo, hi = c - s / 2, c + s / 2
    lo[:, 2] = np.maximum(lo[:, 2], 0.0)
    return torch.tensor(np.concatenate([lo, hi], 1), dtype=torch.float64, device=dev), torch.tensor(cls, device=dev)


def c3_scene(T, shifted, feat_table, dev, H=224, W=224):
    """depth [T,H,W,1] f32, semantic ids [T,H,W,1] i64, instance features [T,56,56,256] f32, poses."""
    boxes, classes = c3_boxes(shifted, dev)
    r = BoxRoomRenderer(H, W, dev, boxes, classes)
    depth, ids, feats, pos, yaw, elev = [], [], [], [], [], []
    for t in range(T):
        p, y, e = r.synthetic.boxroom_pose(t, T)
        d, hit = r.frame(p, y, e)
        depth.append(d[..., None])
        ids.append(torch.where(hit >= 0, classes[hit.clamp(min=0)], torch.zeros_like(hit))[..., None])
        feats.append(feat_table[hit[2::4, 2::4] + 1])
        pos.append(p), yaw.append(y), elev.append(e)
    return dict(position=np.stack(pos), yaw=np.array(yaw, np.float32), elevation=np.array(elev, np.float32),
                depth=torch.stack(depth), semantic=torch.stack(ids), features=torch.stack(feats).contiguous())


def c3_layers(dev):
    from mass_b200.nn.applications.occupancy_projection_layer import OccupancyProjectionLayer
    from mass_b200.nn.applications.resnet_projection_layer import ResNetProjectionLayer
    from mass_b200.nn.applications.semantic_projection_layer import SemanticProjectionLayer
    from mass_b200.utils import synthetic
    kw = dict(camera_height=224, camera_width=224, exact=False, **C3_MAP, **synthetic.MAP_ORIGIN)
    occ = OccupancyProjectionLayer(feature_size=1, **kw).to(dev)
    sems = [SemanticProjectionLayer(feature_size=54, **kw).to(dev) for _ in range(2)]
    ress = [ResNetProjectionLayer(feature_size=256, **kw).to(dev) for _ in range(2)]
    return occ, sems, ress


def c3_match_loop(sems, ress, psd=None):
    """The agent's loop (agent.py:424-450): ask for the next differing class until there is none."""
    if psd is None:
        from mass_b200.utils.experimentation import predict_scene_differences as psd
    moved, calls, pairs, trace = set(), 0, 0, []
    while True:
        obj, g0, g1 = psd(sems[0], sems[1], ress[0], ress[1], moved, list(range(54)), **C3_MATCH_KW)[:3]
        calls += 1
        if obj is None:
            break
        pairs += len(g0)
        trace.append((obj, g0, g1))
        moved.add(obj)
    return calls, pairs, trace


def run_c3(args, R):
    from mass_b200 import _lib
    from mass_b200.utils import synthetic
    dev, world, rank = R.dev, R.world, R.rank
    if args.impl == "reference":
        return run_c3_reference(args, R)
    L = _lib.lib()
    feat_table = torch.rand(C3_OBJECTS + 1, 256, device=dev, generator=torch.Generator(device=dev).manual_seed(7 + rank))
    scenes = [c3_scene(T, shifted, feat_table, dev) for shifted, T in zip((False, True), C3_FRAMES)]
    occ, sems, ress = c3_layers(dev)
    origin = {k: synthetic.MAP_ORIGIN[k] for k in ("origin_y", "origin_x", "origin_z")}
    frames = sum(C3_FRAMES)

    def episode(obs_pair):
        for L_ in (occ, sems[0], sems[1], ress[0], ress[1]):
            L_.reset(**origin)
        for i, obs in enumerate(obs_pair):
            if i == 1:
                occ.reset(**origin)                              # the agent re-maps occupancy per phase
            occ.update_batch(obs)
            sems[i].update_batch(obs)
            ress[i].update_batch(obs)
        return c3_match_loop(sems, ress)

    ev = lambda: torch.cuda.Event(enable_timing=True)            # noqa: E731
    for _ in range(max(args.warmup, 1)):
        episode(scenes)
    # stage split of one untimed episode (host clock around synchronised stages)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for L_ in (occ, sems[0], sems[1], ress[0], ress[1]):
        L_.reset(**origin)
    for i, obs in enumerate(scenes):
        if i == 1:
            occ.reset(**origin)
        occ.update_batch(obs), sems[i].update_batch(obs), ress[i].update_batch(obs)
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    calls, pairs, trace = c3_match_loop(sems, ress)
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    launches0 = L.mb_launch_count()
    import bench
    ms_total = bench.timed_steps(R, lambda: episode(scenes), args.steps)
    launches = int(L.mb_launch_count() - launches0)
    value = world * frames * args.steps / (ms_total * 1e-3)

    # end to end: the same episode from pinned HOST buffers (H2D inside), result read back (D2H)
    e2e = None
    if not args.no_e2e:
        host = [{k: (v.cpu().pin_memory() if torch.is_tensor(v) else v) for k, v in sc.items()} for sc in scenes]
        h2d = sum(int(sc[k].numel() * sc[k].element_size()) * (3 if k == "depth" else 1)
                  for sc in host for k in ("depth", "semantic", "features")) + frames * 48 * 3

        def episode_host():
            calls_, pairs_, _ = episode(host)
            return calls_ + pairs_ + int((occ.data != 0).sum().item())
        episode_host()
        ems = bench.timed_steps(R, episode_host, max(1, min(args.steps, 2)))
        e2e = {"value": world * frames * max(1, min(args.steps, 2)) / (ems * 1e-3), "unit": "frames/s",
               "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 8,
               "note": "depth crosses PCIe once per map it feeds (three layer.update_batch calls per pass, as agent.py "
                       "drives three layers from one observation)"}

    check = None
    if args.check and rank == 0:
        check = c3_check(scenes, sems, ress, trace, dev)
    cpu = None
    if not args.no_cpu_baseline and world == 1 and rank == 0:
        cpu = c3_reference_sample(scenes, sems, ress)
    if rank != 0:
        return
    out = {"metric": "RGB-D frames/sec fused into semantic voxel map", "value": value, "unit": "frames/s",
           "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_total / args.steps,
           "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": c3_workload(), "e2e": e2e, "gpu_launches": launches,
           "stages_ms": {"five_maps": (t1 - t0) * 1e3, "match_loop": (t2 - t1) * 1e3},
           "match": {"predict_scene_differences_calls": calls, "classes_differing": len(trace),
                     "instance_pairs": pairs, "ms_per_call": (t2 - t1) * 1e3 / max(calls, 1)},
           "cpu_baseline": cpu, "check": check}
    print(json.dumps(out))


def c3_workload():
    return {"workload": "c3: episode pair -- walkthrough (250 frames) + unshuffle (500 frames) of a furnished box-room "
                        "(~200 object boxes): occupancy F=1, 2 x semantic F=54 from class ids, 2 x 256-d instance "
                        "features at 56x56, maps 384x384x96 at 0.05 m, then the agent's predict_scene_differences loop "
                        "(find + pairwise L2 + assignment per class) until no class differs",
            "frames_per_step": sum(C3_FRAMES), "episodes": "one per GPU"}


def _oracle_layers_from(sems, ress):
    """OracleLayer shells holding copies of the GPU maps (find / match are pure functions of the maps)."""
    from oracle import oracle
    from mass_b200.utils import synthetic
    out = []
    for group, F in ((sems, 54), (ress, 256)):
        for L_ in group:
            o = oracle.OracleLayer.__new__(oracle.OracleLayer)
            o.camera_height, o.camera_width = L_.camera_height, L_.camera_width
            o.map_height, o.map_width, o.map_depth = L_.map_height, L_.map_width, L_.map_depth
            o.feature_size, o.grid_resolution = F, L_.grid_resolution
            o.bins_x, o.bins_y, o.bins_z = (b.cpu().numpy() for b in (L_.bins_x, L_.bins_y, L_.bins_z))
            o.data = L_.data.cpu().numpy()
            out.append(o)
    return out[:2], out[2:]


def c3_check(scenes, sems, ress, trace, dev):
    """Against the CPU oracle: (1) the first 6 frames of the unshuffle pass through all three kinds of map (occupancy
    bit-exact, values 1e-5); (2) the match stage on the GPU-built maps copied to the host: every class the agent's
    loop returned, with goals in the same order within 1e-5 (find + L2 + assignment indices as the reference's)."""
    from oracle import oracle
    from mass_b200.utils import synthetic
    res = {}
    kw = dict(**C3_MAP, **synthetic.MAP_ORIGIN)
    obs = scenes[1]
    n = 6
    kinds = {"semantic": (224, 54), "features": (56, 256)}
    for name, (cam, F) in kinds.items():
        ref = oracle.OracleLayer(camera_height=cam, camera_width=cam, feature_size=F, nthreads=os.cpu_count() or 8, **kw)
        for t in range(n):
            depth = obs["depth"][t].cpu().numpy()
            if name == "semantic":
                f = np.eye(54, dtype=np.float32)[obs["semantic"][t, ..., 0].cpu().numpy()]
            else:
                f, depth = obs["features"][t].cpu().numpy(), depth[2::4, 2::4]
            ref.update(dict(position=obs["position"][t], yaw=obs["yaw"][t], elevation=obs["elevation"][t], depth=depth,
                            features=f))
        occ_, sems_, ress_ = c3_layers(dev)
        gpu = sems_[0] if name == "semantic" else ress_[0]
        gpu.update_batch({k: v[:n] for k, v in obs.items()})
        g = gpu.data.cpu().numpy()
        res[name + "_map_6_frames"] = {
            "occupancy_bit_exact": bool(np.array_equal((g != 0).any(-1), (ref.data != 0).any(-1))),
            "values_within_1e-5": bool((np.abs(g.astype(np.float64) - ref.data) <= 1e-5 * np.abs(ref.data)).all())}
        del occ_, sems_, ress_, gpu, ref
        torch.cuda.empty_cache()
    osem, ores = _oracle_layers_from(sems, ress)
    moved, same_class, goals_ok, n_cmp = set(), True, True, 0
    for obj, g0, g1 in trace[:4]:
        robj, r0, r1, _ = oracle.predict_scene_differences(osem[0], osem[1], ores[0], ores[1], moved, list(range(54)),
                                                           **C3_MATCH_KW)
        same_class = same_class and (robj == obj) and len(r0) == len(g0)
        if robj == obj and len(r0) == len(g0):
            a0, a1 = torch.stack(g0).cpu().numpy(), torch.stack(g1).cpu().numpy()
            goals_ok = goals_ok and bool(np.allclose(a0, np.stack(r0), rtol=1e-5, atol=1e-5)) and \
                bool(np.allclose(a1, np.stack(r1), rtol=1e-5, atol=1e-5))
        n_cmp += 1
        moved.add(obj)
    res["match_first_%d_classes" % n_cmp] = {"same_class_and_pair_count": same_class, "goals_within_1e-5": goals_ok}
    return res


def c3_reference_sample(scenes, sems, ress):
    """The reference's own CPU path on a bounded sample of the episode: its mapping on 4 frames per kind of map, its
    predict_scene_differences on the GPU-built maps (copied into reference layers) restricted to the first classes
    until one differs -- extrapolated to the episode and said so."""
    from oracle import reference
    from mass_b200.utils import synthetic
    if not reference.available():
        return None
    R_ = reference.load()
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    kw = dict(**C3_MAP, **synthetic.MAP_ORIGIN)
    obs = scenes[1]
    per_frame = {}
    for name, cam, F in (("occupancy", 224, 1), ("semantic", 224, 54), ("features", 56, 256)):
        layer = R_.base.BaseProjectionLayer(camera_height=cam, camera_width=cam, feature_size=F, **kw)
        ts = []
        for t in (0, 100, 200, 300, 400):
            depth = obs["depth"][t].cpu().numpy()
            if name == "occupancy":
                f = np.ones_like(depth)
            elif name == "semantic":
                f = np.eye(54, dtype=np.float32)[obs["semantic"][t, ..., 0].cpu().numpy()]
            else:
                f, depth = obs["features"][t].cpu().numpy(), depth[2::4, 2::4]
            o = dict(position=obs["position"][t], yaw=obs["yaw"][t], elevation=obs["elevation"][t], depth=depth, features=f)
            t0 = time.perf_counter()
            layer.update(o)
            ts.append(time.perf_counter() - t0)
        per_frame[name] = float(np.mean(ts[1:]))
        del layer
    map_s = sum(C3_FRAMES) * (per_frame["semantic"] + per_frame["features"]) + sum(C3_FRAMES) * per_frame["occupancy"]
    # matching: the reference's find on reference layers that hold the GPU-built maps; one class per map timed
    sem_ref = [R_.semantic.SemanticProjectionLayer(camera_height=224, camera_width=224, feature_size=54,
                                                   class_to_colors=torch.zeros(54, 3), **kw) for _ in range(2)]
    res_ref = [R_.base.BaseProjectionLayer(camera_height=56, camera_width=56, feature_size=256, **kw) for _ in range(2)]
    for a, b in zip(sem_ref + res_ref, list(sems) + list(ress)):
        a.data = b.data.cpu()
    t0 = time.perf_counter()
    found = sem_ref[0].find(7, feature_map=res_ref[0], confidence_threshold=0.0, contour_padding=0, contour_threshold=0.0)
    find_s = time.perf_counter() - t0
    # one predict_scene_differences call scans classes (two find() each) until one differs: the GPU loop's trace says
    # how many classes each call had to scan
    from mass_b200.utils.experimentation import ID_TO_OPENABLE, ID_TO_PICKABLE
    calls, _, trace = c3_match_loop(sems, ress)
    scanned, moved = 0, set()
    for obj, _, _ in trace:
        scanned += sum(1 for k in range(obj + 1) if k not in moved and (ID_TO_PICKABLE[k] or ID_TO_OPENABLE[k]))
        moved.add(obj)
    scanned += sum(1 for k in range(54) if k not in moved and (ID_TO_PICKABLE[k] or ID_TO_OPENABLE[k]))
    match_s = scanned * 2 * find_s
    value = sum(C3_FRAMES) / (map_s + match_s)
    return {"value": value, "unit": "frames/s", "cores": cores, "kind": "reference",
            "sample": "EXTRAPOLATED from: the reference's BaseProjectionLayer.update on 4 frames per kind of map "
                      "(%.3f / %.3f / %.3f s per frame for occupancy / semantic one-hot / 256-d at 56x56) and ONE "
                      "SemanticProjectionLayer.find call of the reference on the GPU-built maps (%.2f s, %d instances); the "
                      "agent's loop makes %d predict_scene_differences calls that scan %d classes with two find() each"
                      % (per_frame["occupancy"], per_frame["semantic"], per_frame["features"], find_s, len(found[0]),
                         calls, scanned),
            "episode_seconds_extrapolated": map_s + match_s, "mapping_seconds": map_s, "matching_seconds": match_s}


def run_c3_reference(args, R):
    """--impl reference --config c3: the bounded reference sample alone (needs one GPU pass to build the maps the
    reference's find() is timed on; the timed work is the reference's, on the host cores)."""
    if R.rank != 0:
        return
    dev = R.dev
    feat_table = torch.rand(C3_OBJECTS + 1, 256, device=dev, generator=torch.Generator(device=dev).manual_seed(7))
    scenes = [c3_scene(T, shifted, feat_table, dev) for shifted, T in zip((False, True), C3_FRAMES)]
    occ, sems, ress = c3_layers(dev)
    for i, obs in enumerate(scenes):
        sems[i].update_batch(obs), ress[i].update_batch(obs)
    t0 = time.perf_counter()
    cpu = c3_reference_sample(scenes, sems, ress)
    elapsed = time.perf_counter() - t0
    if cpu is None:
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref not installed"}))
        return
    print(json.dumps({"impl": "reference", "metric": "RGB-D frames/sec fused into semantic voxel map", "value": cpu["value"],
                      "unit": "frames/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
                      "ms_per_step": elapsed * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                      "dtype": "f32", "data": "synthetic", "config": c3_workload(), "cpu_baseline": cpu,
                      "e2e": {"value": cpu["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                      "gpu_launches": 0}))


def main(args):
    import bench
    R = bench.Ranks()
    try:
        if args.config == "c4":
            run_c4(args, R)
        else:
            run_c3(args, R)
    finally:
        R.close()
