"""GPU parity of the instance extraction (find) and matching kernels, through the C ABI, against
the golden fixtures generated from the live reference (tests/golden/make_golden.py) and the CPU oracle."""
import numpy as np
import pytest
import torch

from conftest import golden, golden_kwargs

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return torch.device("cuda:0")


def layer_kwargs(kw):
    keep = ("camera_height", "camera_width", "vertical_fov", "map_height", "map_width", "map_depth",
            "origin_y", "origin_x", "origin_z", "grid_resolution", "interpolation_weight")
    return {k: kw[k] for k in keep if k in kw}


@pytest.fixture(scope="module")
def block_layers(dev):
    from golden.make_golden_maps import build_block_maps
    from mass_b200.nn.base_projection_layer import BaseProjectionLayer
    from mass_b200.nn.applications.semantic_projection_layer import SemanticProjectionLayer
    g = golden("find_match.npz")
    S0, S1, S2, F_feat = [int(v) for v in g["dims"]]
    kw = layer_kwargs(golden_kwargs(g))
    layers = []
    for shift in (0, 1):
        sem, feat = build_block_maps(21, S0, S1, S2, 54, F_feat, shift)
        s = SemanticProjectionLayer(feature_size=54, **kw).to(dev)
        f = BaseProjectionLayer(feature_size=F_feat, **kw).to(dev)
        s.data.copy_(torch.from_numpy(sem))
        f.data.copy_(torch.from_numpy(feat))
        layers.append((s, f))
    return g, layers


def stack(ts):
    return torch.stack(list(ts)).cpu().numpy()


@pytest.mark.parametrize("pad", [0, 1])
def test_find_golden(block_layers, pad):
    g, layers = block_layers
    for cls in (3, 7, 12, 20, 45, 50, 9):
        for m, (s, f) in enumerate(layers):
            tag = "p%d_c%d_m%d" % (pad, cls, m)
            conf, coord, size, feats = s.find(cls, confidence_threshold=0.0, contour_padding=pad,
                                              contour_threshold=0.0, feature_map=f)
            assert len(conf) == int(g["n_" + tag]), tag
            assert np.array_equal(np.array(s.boxes, np.int64).reshape(-1, 4), g["boxes_" + tag]), tag
            assert len(feats) == len(conf)
            if len(conf):
                assert conf[0].is_cuda and conf[0].dim() == 0 and coord[0].shape == (3,)
                np.testing.assert_allclose(stack(conf), g["conf_" + tag], rtol=1e-5)
                np.testing.assert_allclose(stack(coord), g["coord_" + tag], rtol=1e-5, atol=1e-6)
                np.testing.assert_allclose(stack(size), g["size_" + tag], rtol=1e-5)
                np.testing.assert_allclose(stack(feats), g["feat_" + tag], rtol=1e-5)
            # without a feature map the reference returns None for the features
            assert s.find(cls, 0.0, pad, 0.0, None)[3] is None


def test_class_presence_vs_oracle(block_layers, oracle):
    from mass_b200.utils import instances
    g, layers = block_layers
    s = layers[0][0]
    data = s.data.cpu().numpy()
    for pad in (0, 1, 2, 3):
        for cls, thr in ((7, 0.0), (20, 0.05), (9, 0.0)):
            img = instances.class_presence(s, cls, pad, thr).cpu().numpy()
            assert np.array_equal(img, oracle.class_presence(data, cls, pad, thr)), (pad, cls)


def test_find_confidence_threshold_and_errors(block_layers):
    g, layers = block_layers
    s, f = layers[0]
    all_conf = stack(s.find(7, 0.0, 0, 0.0, f)[0])
    thr = float(np.sort(all_conf)[len(all_conf) // 2])
    conf = s.find(7, thr, 0, 0.0, f)[0]
    assert len(conf) == int((all_conf > np.float32(thr)).sum())
    with pytest.raises(IndexError):
        s.find(54, 0.0, 0, 0.0, None)
    with pytest.raises(RuntimeError):
        s.find(7, 0.0, 0, 0.0, f.__class__(feature_size=16, **layer_kwargs(golden_kwargs(g))))   # CPU feature map


def test_find_cache_follows_the_map(block_layers, dev):
    """find() results are cached against the map state: kernel updates, torch in-place writes and feature-map
    changes must all invalidate them."""
    g, layers = block_layers
    s, f = layers[1]
    saved_s, saved_f = s.data.clone(), f.data.clone()
    try:
        conf, coord, size, feats = s.find(7, 0.0, 0, 0.0, f)
        assert len(conf) > 0
        again = s.find(7, 0.0, 0, 0.0, f)
        assert torch.equal(torch.stack(again[0]), torch.stack(conf))
        f.data.mul_(2.0)                                           # torch in-place op on the feature map
        feats2 = s.find(7, 0.0, 0, 0.0, f)[3]
        np.testing.assert_allclose(stack(feats2), 2.0 * stack(feats), rtol=1e-6)
        s.data[..., 7].zero_()                                     # torch in-place op on the semantic map
        assert len(s.find(7, 0.0, 0, 0.0, f)[0]) == 0
        # a kernel update (raw-pointer write) also invalidates: put one observation into the map
        obs = dict(position=np.array([0.3, -0.1, 0.9], np.float32), yaw=np.float32(0.3), elevation=np.float32(-0.4),
                   depth=np.full((s.camera_height, s.camera_width, 1), 0.6, np.float32),
                   semantic=np.full((s.camera_height, s.camera_width, 1), 7, np.int64))
        s.update(obs)
        assert len(s.find(7, 0.0, 0, 0.0, f)[0]) > 0
    finally:
        s.data.copy_(saved_s)
        f.data.copy_(saved_f)


def test_find_cache_follows_raw_pointer_writers(block_layers, dev):
    """Writers that bypass update(): the ordered affine apply of a sharded partial bumps the map state itself; a
    CUDA graph the CALLER captured around update_prepared runs no Python on replay, so the caller says
    mark_dirty() -- after which find() sees the new map."""
    from mass_b200.nn import sharded
    g, layers = block_layers
    s, f = layers[1]
    saved = s.data.clone()
    try:
        before = s.find(7, 0.0, 0, 0.0, f)
        assert len(before[0]) > 0
        occ = (s.data[..., 7] != 0).nonzero()
        idx = ((occ[:, 0] * s.data.shape[1] + occ[:, 1]) * s.data.shape[2] + occ[:, 2]).to(torch.int64)
        state = s.map_state()
        sharded.apply_partial(s, idx, torch.zeros(idx.numel(), device=dev),
                              torch.zeros(idx.numel(), s.data.shape[3], device=dev))     # wipes class 7's voxels
        assert s.map_state() != state
        assert len(s.find(7, 0.0, 0, 0.0, f)[0]) == 0
        # caller-captured graph: one frame that paints class 7 back in
        s.exact = False
        H, W = s.camera_height, s.camera_width
        prep = s.prepare_batch(dict(position=np.array([[0.3, -0.1, 0.9]], np.float32), yaw=np.array([0.3], np.float32),
                                    elevation=np.array([-0.4], np.float32),
                                    depth=torch.full((1, H, W, 1), 0.6, device=dev),
                                    class_ids=torch.full((1, H, W), 7, dtype=torch.int64, device=dev)))
        s.update_prepared(prep)
        s.data.copy_(saved)
        sharded.apply_partial(s, idx, torch.zeros(idx.numel(), device=dev), torch.zeros(idx.numel(), s.data.shape[3], device=dev))
        assert len(s.find(7, 0.0, 0, 0.0, f)[0]) == 0
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            s.update_prepared(prep)
        state = s.map_state()
        graph.replay()
        assert s.map_state() == state                         # a replay runs no Python: the layer cannot know
        s.mark_dirty()
        assert len(s.find(7, 0.0, 0, 0.0, f)[0]) > 0
    finally:
        s.exact = True
        s.data.copy_(saved)
        s.mark_dirty()


def test_pairwise_l2_golden(dev):
    from mass_b200.utils import instances
    g = golden("pairwise.npz")
    for a, b, d in ((g["a"], g["b"], g["d"]), (g["a3"], g["b3"], g["d3"])):
        got = instances.pairwise_l2(torch.from_numpy(a).to(dev), torch.from_numpy(b).to(dev)).cpu().numpy()
        np.testing.assert_allclose(got, d, rtol=2e-6)
    empty = instances.pairwise_l2(torch.zeros(0, 8, device=dev), torch.zeros(3, 8, device=dev))
    assert tuple(empty.shape) == (0, 3)


def test_lsap_golden(dev):
    from mass_b200.utils import instances
    g = golden("lsap.npz")
    for c, (nr, nc), r, cc in zip(g["costs"], g["shapes"], g["rows"], g["cols"]):
        rows, cols = instances.linear_sum_assignment(torch.from_numpy(np.ascontiguousarray(c[:nr, :nc])).to(dev))
        k = min(nr, nc)
        assert rows.tolist() == r[:k].tolist() and cols.tolist() == cc[:k].tolist(), (nr, nc)
    for name in ("big", "rect"):
        rows, cols = instances.linear_sum_assignment(torch.from_numpy(g[name]).to(dev))
        assert np.array_equal(rows, g[name + "_rows"]) and np.array_equal(cols, g[name + "_cols"])
    rows, cols = instances.linear_sum_assignment(torch.from_numpy(np.ascontiguousarray(g["rect"].T)).to(dev))
    assert np.array_equal(rows, g["rect_t_rows"]) and np.array_equal(cols, g["rect_t_cols"])


def test_lsap_fuzz_vs_oracle(dev, oracle):
    from mass_b200.utils import instances
    rng = np.random.default_rng(5)
    for k in range(120):
        nr, nc = rng.integers(1, 40, 2)
        c = rng.integers(0, 3, (nr, nc)).astype(np.float64) if k % 2 else rng.random((nr, nc)).astype(np.float32)
        rows, cols = instances.linear_sum_assignment(torch.from_numpy(c).to(dev))
        r, cc = oracle.lsap(c)
        assert rows.tolist() == r.tolist() and cols.tolist() == cc.tolist(), (k, nr, nc)
    c = np.full((3, 3), np.inf)
    with pytest.raises(ValueError):
        instances.linear_sum_assignment(torch.from_numpy(c).to(dev))
    with pytest.raises(ValueError):
        instances.linear_sum_assignment(torch.full((2, 2), float("nan"), device=dev))


def test_predict_scene_differences_golden(block_layers):
    from mass_b200.utils.experimentation import order_goals, predict_scene_differences
    g, ((s0, f0), (s1, f1)) = block_layers
    for k in range(int(g["num_psd"])):
        use_feat = bool(g["psd%d_use_feat" % k])
        obj, g0, g1 = predict_scene_differences(
            s0, s1, f0 if use_feat else None, f1 if use_feat else None,
            set(g["psd%d_moved" % k].tolist()), g["psd%d_cands" % k].tolist(),
            confidence_threshold=0.0, contour_padding=0, contour_threshold=0.0, distance_threshold=0.05)
        assert (-1 if obj is None else obj) == int(g["psd%d_obj" % k]), k
        assert len(g0) == len(g["psd%d_g0" % k])
        if g0:
            np.testing.assert_allclose(stack(g0), g["psd%d_g0" % k], rtol=1e-5, atol=1e-6)
            np.testing.assert_allclose(stack(g1), g["psd%d_g1" % k], rtol=1e-5, atol=1e-6)
            assert order_goals(g0, g1).tolist() == g["psd%d_order" % k].tolist()


def test_c3_scale_matching_vs_oracle(dev, oracle):
    """~200 instances x 256-d features (BASELINE config 3 scale): cost matrix within 2e-6 and the
    assignment identical to the oracle's on the SAME cost matrix, plus equal on its own matrix."""
    from mass_b200.utils import instances
    rng = np.random.default_rng(11)
    a = rng.random((200, 256), dtype=np.float32)
    b = np.concatenate([a[rng.permutation(200)[:190]] + 0.01 * rng.standard_normal((190, 256)).astype(np.float32),
                        rng.random((17, 256), dtype=np.float32)])
    d_gpu = instances.pairwise_l2(torch.from_numpy(a).to(dev), torch.from_numpy(b).to(dev))
    d_ref = oracle.pairwise_l2(a, b)
    np.testing.assert_allclose(d_gpu.cpu().numpy(), d_ref, rtol=2e-6)
    rows, cols = instances.linear_sum_assignment(d_gpu)
    r, c = oracle.lsap(d_gpu.cpu().numpy())
    assert np.array_equal(rows, r) and np.array_equal(cols, c)
    r2, c2 = oracle.lsap(d_ref)
    assert np.array_equal(rows, r2) and np.array_equal(cols, c2)


def test_cosine_best_match(dev):
    """Additional op (not the reference's matcher): against float64 numpy, ties to the first maximum."""
    from mass_b200.utils import instances
    rng = np.random.default_rng(31)
    a = rng.standard_normal((200, 256)).astype(np.float32)
    b = np.concatenate([a[rng.permutation(200)[:150]] * rng.uniform(0.5, 2.0, (150, 1)).astype(np.float32)
                        + 0.05 * rng.standard_normal((150, 256)).astype(np.float32),
                        rng.standard_normal((40, 256)).astype(np.float32)])
    b[7] = b[3]                                       # an exact duplicate: the lower index must win
    b[100] = 0.0                                      # a zero vector never wins against a positive similarity
    best, sim = instances.cosine_best_match(torch.from_numpy(a).to(dev), torch.from_numpy(b).to(dev))
    a64, b64 = a.astype(np.float64), b.astype(np.float64)
    na, nb = np.linalg.norm(a64, axis=1), np.linalg.norm(b64, axis=1)
    den = na[:, None] * nb[None, :]
    ref = np.where(den > 0, (a64 @ b64.T) / np.where(den > 0, den, 1.0), 0.0)
    assert best.cpu().numpy().tolist() == ref.argmax(axis=1).tolist()
    np.testing.assert_allclose(sim.cpu().numpy(), ref.max(axis=1), rtol=1e-6)
    assert 7 not in best.cpu().numpy().tolist() or 3 not in ref.argmax(axis=1).tolist()
    e_best, e_sim = instances.cosine_best_match(torch.zeros(3, 8, device=dev), torch.zeros(0, 8, device=dev))
    assert e_best.tolist() == [-1, -1, -1]


@pytest.mark.parametrize("n,m,d", [(1500, 2300, 256), (300, 1300, 70), (128, 128, 32)])
def test_cosine_best_match_tensor_cores(dev, n, m, d):
    """The tcgen05 path (3 x TF32 ranks, float64 decides) returns the SIMT kernel's indices -- first maximum on ties,
    duplicates and zero rows included -- and both equal float64 numpy."""
    from mass_b200.utils import instances
    rng = np.random.default_rng(41)
    a = rng.standard_normal((n, d)).astype(np.float32)
    b = rng.standard_normal((m, d)).astype(np.float32)
    k = min(n, m) // 2
    b[:k] = a[rng.permutation(n)[:k]] * rng.uniform(0.5, 2.0, (k, 1)).astype(np.float32) \
        + 0.02 * rng.standard_normal((k, d)).astype(np.float32)            # near-matches: close calls between candidates
    b[m - 1] = b[5]                                                          # exact duplicates: the lower index must win
    b[m // 2] = b[9]
    b[17] = 0.0                                                              # zero rows never win a positive similarity
    a[3] = 0.0                                                               # a zero query: similarity 0 everywhere, index 0
    ta, tb = torch.from_numpy(a).to(dev), torch.from_numpy(b).to(dev)
    best_tc, sim_tc = instances.cosine_best_match(ta, tb, tensor_cores=True)
    best_sm, sim_sm = instances.cosine_best_match(ta, tb, tensor_cores=False)
    torch.cuda.synchronize()
    assert torch.equal(best_tc, best_sm)
    np.testing.assert_allclose(sim_tc.cpu().numpy(), sim_sm.cpu().numpy(), rtol=1e-6, atol=1e-7)
    a64, b64 = a.astype(np.float64), b.astype(np.float64)
    den = np.linalg.norm(a64, axis=1)[:, None] * np.linalg.norm(b64, axis=1)[None, :]
    ref = np.where(den > 0, (a64 @ b64.T) / np.where(den > 0, den, 1.0), 0.0)
    assert best_tc.cpu().numpy().tolist() == ref.argmax(axis=1).tolist()
    assert int(best_tc[3]) == 0 and float(sim_tc[3]) == 0.0
