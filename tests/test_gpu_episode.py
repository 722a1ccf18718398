"""BASELINE config 3 in miniature, end to end on the GPU against the CPU oracle: a walkthrough and an unshuffle
pass over a furnished box-room build two semantic maps (class ids -> one-hot) and two instance-feature maps
(quarter-resolution camera), then predict_scene_differences extracts instances per class and matches them."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

H = W = 64
T = 20
F_FEAT = 16
KW = dict(vertical_fov=90.0, map_height=96, map_width=96, map_depth=32, grid_resolution=0.1, interpolation_weight=0.5,
          origin_x=0.0, origin_y=0.0, origin_z=0.9)


def _boxes(shifted):
    rng = np.random.default_rng(17)
    boxes, classes = [], []
    for k in range(10):
        c = rng.uniform([-3.0, -2.2, 0.0], [3.0, 2.2, 0.8])
        s = rng.uniform(0.25, 0.5, 3)
        if shifted and k in (2, 5, 7):
            c[:2] += rng.uniform(0.6, 0.9, 2) * rng.choice([-1, 1], 2)
        boxes.append(np.concatenate([c - s / 2, c + s / 2]))
        classes.append([3, 3, 7, 7, 7, 12, 20, 45, 45, 50][k])
    boxes = np.array(boxes)
    boxes[:, 2] = np.maximum(boxes[:, 2], 0.0)
    return boxes, np.array(classes)


def _frames(shifted):
    from mass_b200.utils import synthetic
    boxes, classes = _boxes(shifted)
    rays = synthetic.camera_rays(H, W)
    feat_table = np.random.default_rng(5).random((len(boxes) + 1, F_FEAT)).astype(np.float32)
    out = []
    for t in range(T):
        pos, yaw, elev = synthetic.boxroom_pose(t, T)
        depth, hit = synthetic.render_depth(rays, pos, yaw, elev, boxes)
        ids = np.where(hit >= 0, classes[np.maximum(hit, 0)], 0).astype(np.int64)
        feats = feat_table[hit[2::4, 2::4] + 1] * (1.0 + 0.01 * np.random.default_rng(100 + t).random((H // 4, W // 4, 1)).astype(np.float32))
        out.append(dict(position=pos, yaw=yaw, elevation=elev, depth=depth[..., None], semantic=ids[..., None],
                        features=np.ascontiguousarray(feats, dtype=np.float32)))
    return out


def _oracle_maps(oracle, frames):
    sem = oracle.OracleLayer(camera_height=H, camera_width=W, feature_size=54, **KW)
    feat = oracle.OracleLayer(camera_height=H // 4, camera_width=W // 4, feature_size=F_FEAT, **KW)
    for f in frames:
        onehot = np.eye(54, dtype=np.float32)[f["semantic"][..., 0]]
        sem.update(dict(position=f["position"], yaw=f["yaw"], elevation=f["elevation"], depth=f["depth"], features=onehot))
        feat.update(dict(position=f["position"], yaw=f["yaw"], elevation=f["elevation"], depth=f["depth"][2::4, 2::4],
                         features=f["features"]))
    return sem, feat


def _gpu_maps(frames, exact, dev):
    from mass_b200.nn.applications.resnet_projection_layer import ResNetProjectionLayer
    from mass_b200.nn.applications.semantic_projection_layer import SemanticProjectionLayer
    sem = SemanticProjectionLayer(camera_height=H, camera_width=W, feature_size=54, exact=exact, **KW).to(dev)
    feat = ResNetProjectionLayer(camera_height=H, camera_width=W, feature_size=F_FEAT, exact=exact, **KW).to(dev)
    if exact:
        for f in frames:
            sem.update(f)
            feat.update(f)
    else:
        sem.update_batch(frames)
        feat.update_batch(frames)
    return sem, feat


@pytest.mark.parametrize("exact", [True, False])
def test_episode_pair_maps_find_and_match(oracle, exact):
    from mass_b200.utils.experimentation import predict_scene_differences
    dev = torch.device("cuda:0")
    walk, unshuffle = _frames(False), _frames(True)
    o0, of0 = _oracle_maps(oracle, walk)
    o1, of1 = _oracle_maps(oracle, unshuffle)
    g0, gf0 = _gpu_maps(walk, exact, dev)
    g1, gf1 = _gpu_maps(unshuffle, exact, dev)
    for got, ref in ((g0, o0), (gf0, of0), (g1, o1), (gf1, of1)):
        a, b = got.data.cpu().numpy(), ref.data
        assert np.array_equal((a != 0).any(-1), (b != 0).any(-1))
        if exact:
            assert np.array_equal(a, b)
        else:
            assert (np.abs(a.astype(np.float64) - b) <= 1e-5 * np.abs(b)).all()
    kw = dict(confidence_threshold=0.0, contour_padding=0, contour_threshold=0.0, distance_threshold=0.05)
    moved = set()
    found_any = False
    for _ in range(4):                                  # the agent's loop: agent.py:424-450
        obj, a0, a1 = predict_scene_differences(g0, g1, gf0, gf1, moved, list(range(54)), **kw)
        robj, r0, r1, _ = oracle.predict_scene_differences(o0, o1, of0, of1, moved, list(range(54)), **kw)
        assert obj == robj
        if obj is None:
            break
        found_any = True
        assert len(a0) == len(r0) and len(a1) == len(r1)
        np.testing.assert_allclose(torch.stack(a0).cpu().numpy(), np.stack(r0), rtol=1e-5, atol=1e-5)
        np.testing.assert_allclose(torch.stack(a1).cpu().numpy(), np.stack(r1), rtol=1e-5, atol=1e-5)
        moved.add(obj)
    assert found_any
    # per class: same instances (count, order, boxes) and pooled statistics in both maps
    for cls in (3, 7, 12, 20, 45, 50):
        for gs, gf, os_, of in ((g0, gf0, o0, of0), (g1, gf1, o1, of1)):
            conf, coord, size, feats = gs.find(cls, 0.0, 0, 0.0, gf)
            rconf, rcoord, rsize, rfeats, rboxes = oracle.find(os_, cls, 0.0, 0, 0.0, of)
            assert [tuple(b) for b in gs.boxes] == [tuple(b) for b in rboxes]
            if rconf:
                np.testing.assert_allclose(torch.stack(conf).cpu().numpy(), np.stack(rconf), rtol=1e-5)
                np.testing.assert_allclose(torch.stack(size).cpu().numpy(), np.stack(rsize), rtol=1e-5)
                np.testing.assert_allclose(torch.stack(feats).cpu().numpy(), np.stack(rfeats), rtol=1e-5, atol=1e-6)
