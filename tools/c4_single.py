"""BASELINE config 4 at FULL map size on one GPU: 720x1280 depth + 54-class probabilities fused into a 960x960x240 map
at 0.02 m (47.8 GB of map in HBM), 32 resident frames per call (ring of <= 32 frames per GPU).
Prints frames/s of the batched update and the fraction of the HBM roofline on the survey's bytes per frame
(B_frame = 4hw + 4hwF + 8F U_f + 48 with U_f counted here from the touched-voxel list of a one-frame call).
    python tools/c4_single.py [frames_per_call] [calls]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    from mass_b200.nn.base_projection_layer import BaseProjectionLayer
    from mass_b200.utils import synthetic

    dev = torch.device("cuda", 0)
    per_call = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    calls = int(sys.argv[2]) if len(sys.argv) > 2 else 4
    H, W, F = 720, 1280, 54
    kw = dict(camera_height=H, camera_width=W, vertical_fov=90.0, map_height=960, map_width=960, map_depth=240,
              feature_size=F, grid_resolution=0.02, interpolation_weight=0.5, **synthetic.MAP_ORIGIN)
    rays = synthetic.camera_rays(H, W)
    layer = BaseProjectionLayer(exact=False, **kw).to(dev)
    print("map %.1f GB" % (layer.data.numel() * 4 / 1e9), flush=True)

    def frames(t0, n):
        pos, yaw, elev, depth = [], [], [], []
        for t in range(t0, t0 + n):
            p, y, e = synthetic.boxroom_pose(t, 4096)         # config 4: 4096 frames around the room
            d, _ = synthetic.render_depth(rays, p, y, e)
            pos.append(p), yaw.append(y), elev.append(e), depth.append(d[..., None])
        g = torch.Generator(device=dev).manual_seed(1000 + t0)
        low = torch.softmax(4 * torch.randn(n, H // 8, W // 8, F, device=dev, generator=g), dim=-1)
        probs = low.repeat_interleave(8, 1).repeat_interleave(8, 2).contiguous()
        return dict(position=np.stack(pos), yaw=np.array(yaw, np.float32), elevation=np.array(elev, np.float32),
                    depth=torch.from_numpy(np.stack(depth)).to(dev), features=probs)

    # touched voxels of single frames (the survey's U_f): occupancy of a fresh map after one frame
    one = frames(0, 1)
    layer.update_batch(one)
    touched = int((layer.data != 0).any(-1).sum().item())
    layer.data.zero_()
    b_frame = 4 * H * W + 4 * H * W * F + 8 * F * touched + 48
    print("touched voxels of frame 0: %d -> B_frame %.1f MB" % (touched, b_frame / 1e6), flush=True)

    batches = [layer.prepare_batch(frames(c * per_call, per_call)) for c in range(2)]     # two resident rings
    for w in range(2):
        layer.update_prepared(batches[w % 2])
    layer.check()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for c in range(calls):
        layer.update_prepared(batches[c % 2])
    ev1.record()
    layer.check()
    ms = ev0.elapsed_time(ev1) / calls
    fps = per_call / ms * 1e3
    peak = 6453.1
    try:
        import json
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    print("c4 single GPU: %d frames per call, %.2f ms per call, %.0f frames/s, %.1f GB/s algorithmic = %.3f of %.0f GB/s"
          % (per_call, ms, fps, fps * b_frame / 1e9, fps * b_frame / 1e9 / peak, peak))
    print("peak memory %.1f GB" % (torch.cuda.max_memory_allocated() / 1e9))


if __name__ == "__main__":
    main()
