"""Workload statistics of the C2 box-room walkthrough by voxel brick (design aid, CPU only):
touched voxels per frame, entries per brick, critical path of the heaviest brick."""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import oracle
from mass_b200.utils import synthetic

T = 500
step = int(sys.argv[1]) if len(sys.argv) > 1 else 5
kw = dict(camera_height=224, camera_width=224, vertical_fov=90.0, map_height=384, map_width=384,
          map_depth=96, feature_size=1, grid_resolution=0.05, interpolation_weight=0.5, **synthetic.MAP_ORIGIN)
L = oracle.OracleLayer(**kw)
rays = synthetic.camera_rays(224, 224)
S = (384, 384, 96)
for bs in ((8, 8, 8), (8, 8, 4), (4, 4, 4), (4, 4, 8)):
    nb = tuple((s + b - 1) // b for s, b in zip(S, bs))
    ent = np.zeros(nb, np.int64)      # pixel entries per brick (footprint intersects)
    con = np.zeros(nb, np.int64)      # contributions per brick
    frames_per_brick = np.zeros(nb, np.int64)
    touched = []
    tot_entries = 0
    maxgroup = 0
    for t in range(0, T, step):
        pos, yaw, el = synthetic.boxroom_pose(t, T)
        d, _ = synthetic.render_depth(rays, pos, yaw, el)
        eye, up = oracle.eye_up(yaw, el)
        ix, iy, iz, rx, ry, rz, pix = oracle.bin_rays(L.bins_x, L.bins_y, L.bins_z, pos, oracle.transform_rays(L.rays, eye, up), d)
        i0, i1, i2, r0, r1, r2 = iy, ix, iz, ry, rx, rz
        lo, hi = [], []
        for i, r, s in ((i0, r0, S[0]), (i1, r1, S[1]), (i2, r2, S[2])):
            l = np.where(r < .5, np.maximum(i - 1, 0), i)
            h = np.where(r < .5, i, np.minimum(i + 1, s - 1))
            lo.append(l); hi.append(h)
        vox = set()
        keys = []
        bk = []
        for s in range(8):
            c = [(hi if (s >> (2 - a)) & 1 else lo)[a] for a in range(3)]
            keys.append((c[0] * S[1] + c[1]) * S[2] + c[2])
            bk.append(((c[0] // bs[0]) * nb[1] + c[1] // bs[1]) * nb[2] + c[2] // bs[2])
        keys = np.stack(keys); bk = np.stack(bk)          # [8, N]
        touched.append(np.unique(keys).size)
        np.add.at(con.reshape(-1), bk.reshape(-1), 1)
        # entries: distinct bricks per pixel
        sb = np.sort(bk, axis=0)
        first = np.ones_like(sb, bool); first[1:] = sb[1:] != sb[:-1]
        eb = sb[first]
        tot_entries += eb.size
        cnt = np.bincount(eb, minlength=ent.size)
        maxgroup = max(maxgroup, cnt.max())
        ent.reshape(-1)[:] += cnt
        frames_per_brick.reshape(-1)[:] += cnt > 0
    nfr = len(range(0, T, step))
    scale = T / nfr
    nz = ent > 0
    print("brick %s: touched bricks %d, entries/frame %.0f (dup %.2f), contributions/frame %.0f, touched vox/frame %.0f"
          % (bs, nz.sum(), tot_entries / nfr, tot_entries / nfr / 50176, con.sum() / nfr, np.mean(touched)))
    e = ent[nz] * scale
    print("   entries per brick over episode: mean %.0f  p50 %.0f  p99 %.0f  max %.0f ; total/148 = %.0f ; max/(total/148) = %.2f"
          % (e.mean(), np.median(e), np.percentile(e, 99), e.max(), e.sum() / 148, e.max() / (e.sum() / 148)))
    print("   max entries in one (brick,frame) group %d ; frames per brick: mean %.1f max %.0f ; groups total %.0f (per frame %.0f)"
          % (maxgroup, frames_per_brick[nz].mean() * scale, frames_per_brick.max() * scale, frames_per_brick.sum() * scale, frames_per_brick.sum() / nfr))
