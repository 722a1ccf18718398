"""Pins the CPU oracle (oracle/) to golden vectors produced by the unmodified
reference (tests/golden/make_golden.py).  CPU only."""
import hashlib

import numpy as np

from conftest import golden, golden_kwargs


def sha(a):
    return np.frombuffer(hashlib.sha256(np.ascontiguousarray(a).tobytes()).digest(), np.uint8)


def layer_kwargs(kw):
    kw = dict(kw)
    kw.pop("class_to_colors", None)
    return kw


def test_kat_tiny(oracle):
    g = golden("kat_tiny.npz")
    L = oracle.OracleLayer(camera_height=2, camera_width=2, vertical_fov=90.0, map_height=8, map_width=8,
                           map_depth=4, feature_size=2, grid_resolution=0.5, interpolation_weight=0.5)
    assert np.array_equal(L.rays, g["rays"])
    for b in ("bins_x", "bins_y", "bins_z"):
        assert np.array_equal(getattr(L, b), g[b])
    # SURVEY.md section 4 known answers
    assert np.array_equal(L.rays, np.array([[[-.5, .5, -1], [.5, .5, -1]], [[-.5, -.5, -1], [.5, -.5, -1]]], np.float32))
    eye, up = oracle.eye_up(g["obs_yaw"], g["obs_elevation"])
    oriented = oracle.transform_rays(L.rays, eye, up)
    out = oracle.bin_rays(L.bins_x, L.bins_y, L.bins_z, g["obs_position"], oriented, g["obs_depth"])
    for name, a in zip(("ind_x", "ind_y", "ind_z", "ratio_x", "ratio_y", "ratio_z"), out):
        assert np.array_equal(a, g[name]), name
    assert out[0].tolist() == [4] and out[1].tolist() == [3] and out[2].tolist() == [3]
    obs = {k[4:]: g[k] for k in g.files if k.startswith("obs_")}
    L.update(obs)
    assert np.array_equal(L.data, g["data1"])
    assert abs(L.data[3, 4, 3, 0] - 0.144) < 1e-6 and int((L.data[..., 0] != 0).sum()) == 8
    L.update(obs)
    assert np.array_equal(L.data, g["data2"])


def test_pose_rotation(oracle):
    g = golden("pose.npz")
    eye, up = oracle.eye_up(g["yaw"], g["elevation"])          # batched ATen ops
    assert np.array_equal(eye, g["eye"]) and np.array_equal(up, g["up"])
    for k in range(len(g["yaw"])):
        assert np.array_equal(oracle.rotation_from_eye_up(g["eye"][k], g["up"][k]), g["rot"][k]), k
    assert np.array_equal(oracle.transform_rays(g["rays"], g["eye"][9], g["up"][9]), g["oriented9"])


def _run_seq(oracle, g, T, with_bins=True, nthreads=1):
    kw = layer_kwargs(golden_kwargs(g))
    L = oracle.OracleLayer(nthreads=nthreads, **kw)
    for t in range(T):
        obs = {k: g[k][t] for k in ("position", "yaw", "elevation", "depth", "features")}
        if with_bins:
            eye, up = oracle.eye_up(obs["yaw"], obs["elevation"])
            out = oracle.bin_rays(L.bins_x, L.bins_y, L.bins_z, obs["position"],
                                  oracle.transform_rays(L.rays, eye, up), obs["depth"])
            for name, a in zip(("ind_x", "ind_y", "ind_z", "ratio_x", "ratio_y", "ratio_z"), out):
                assert np.array_equal(a, g["%s_%d" % (name, t)]), (name, t)
        L.update(obs)
        if "data_%d" % t in g.files:
            assert np.array_equal(L.data, g["data_%d" % t]), t
    return L


def test_small_sequence_bitwise(oracle):
    _run_seq(oracle, golden("small_seq.npz"), 6)


def test_small_sequence_threads(oracle):
    _run_seq(oracle, golden("small_seq.npz"), 6, with_bins=False, nthreads=3)


def test_border_and_special_depths(oracle):
    _run_seq(oracle, golden("border.npz"), 2)


def test_lowres_feature_upsampling(oracle):
    g = golden("lowres.npz")
    L = _run_seq(oracle, g, 3, with_bins=False)
    assert np.array_equal(L.data, g["data"])


def test_c1_frames(oracle):
    """BASELINE config 1 (224x224, 54 classes, 384x384x96 @ 0.05 m): digests of
    the reference's indices / ratios / occupancy / touched rows."""
    from mass_b200.utils import synthetic
    g = golden("c1_frames.npz")
    L = oracle.OracleLayer(nthreads=4, **layer_kwargs(golden_kwargs(g)))
    for n in range(2):
        obs = dict(position=g["position_%d" % n], yaw=g["yaw_%d" % n], elevation=g["elevation_%d" % n],
                   depth=g["depth_%d" % n][..., None], features=synthetic.upsample(g["probs_low_%d" % n], 8))
        eye, up = oracle.eye_up(obs["yaw"], obs["elevation"])
        out = oracle.bin_rays(L.bins_x, L.bins_y, L.bins_z, obs["position"],
                              oracle.transform_rays(L.rays, eye, up), obs["depth"])
        assert out[0].size == int(g["n_valid_%d" % n])
        for name, a in zip(("ind_x", "ind_y", "ind_z", "ratio_x", "ratio_y", "ratio_z"), out[:6]):
            a = a.astype(np.int32) if a.dtype == np.int64 else a
            assert np.array_equal(sha(a), g["%s_sha_%d" % (name, n)]), (name, n)
        L.update(obs)
        data = L.data.reshape(-1, 54)
        occ = np.flatnonzero((data != 0).any(-1))
        assert occ.size == int(g["occ_count_%d" % n])
        assert np.array_equal(sha(occ.astype(np.int64)), g["occ_sha_%d" % n])
        assert np.array_equal(data[g["sample_idx_%d" % n]], g["sample_rows_%d" % n])
        assert np.array_equal(sha(data[occ]), g["rows_sha_%d" % n])


def test_lsap_matches_scipy_golden(oracle):
    g = golden("lsap.npz")
    for c, (nr, nc), r, cc in zip(g["costs"], g["shapes"], g["rows"], g["cols"]):
        rows, cols = oracle.lsap(c[:nr, :nc])
        k = min(nr, nc)
        assert rows.tolist() == r[:k].tolist() and cols.tolist() == cc[:k].tolist(), (nr, nc)
    rows, cols = oracle.lsap(g["big"])
    assert np.array_equal(rows, g["big_rows"]) and np.array_equal(cols, g["big_cols"])
    rows, cols = oracle.lsap(g["rect"])
    assert np.array_equal(rows, g["rect_rows"]) and np.array_equal(cols, g["rect_cols"])
    rows, cols = oracle.lsap(g["rect"].T)
    assert np.array_equal(rows, g["rect_t_rows"]) and np.array_equal(cols, g["rect_t_cols"])


def test_lsap_matches_installed_scipy(oracle):
    from scipy.optimize import linear_sum_assignment
    rng = np.random.default_rng(77)
    for k in range(300):
        nr, nc = rng.integers(1, 12, 2)
        c = rng.integers(0, 3, (nr, nc)).astype(float) if k % 2 else rng.random((nr, nc))
        r, cc = linear_sum_assignment(c)
        rows, cols = oracle.lsap(c)
        assert rows.tolist() == r.tolist() and cols.tolist() == cc.tolist()


def test_pairwise_l2(oracle):
    g = golden("pairwise.npz")
    np.testing.assert_allclose(oracle.pairwise_l2(g["a"], g["b"]), g["d"], rtol=2e-6)
    np.testing.assert_allclose(oracle.pairwise_l2(g["a3"], g["b3"]), g["d3"], rtol=2e-6)


def _block_layers(oracle, g):
    from golden.make_golden_maps import build_block_maps
    S0, S1, S2, F_feat = [int(v) for v in g["dims"]]
    kw = layer_kwargs(golden_kwargs(g))
    layers = []
    for shift in (0, 1):
        sem, feat = build_block_maps(21, S0, S1, S2, 54, F_feat, shift)
        s = oracle.OracleLayer(feature_size=54, **kw)
        f = oracle.OracleLayer(feature_size=F_feat, **kw)
        s.data[...] = sem
        f.data[...] = feat
        layers.append((s, f))
    return layers


def test_find_instances(oracle):
    g = golden("find_match.npz")
    layers = _block_layers(oracle, g)
    for pad in (0, 1):
        for cls in (3, 7, 12, 20, 45, 50, 9):
            for m, (s, f) in enumerate(layers):
                tag = "p%d_c%d_m%d" % (pad, cls, m)
                conf, coord, size, feats, boxes = oracle.find(s, cls, confidence_threshold=0.0,
                                                              contour_padding=pad, contour_threshold=0.0,
                                                              feature_map=f)
                assert len(conf) == int(g["n_" + tag]), tag
                assert np.array_equal(np.array(boxes, np.int64).reshape(-1, 4), g["boxes_" + tag]), tag
                if len(conf):
                    np.testing.assert_allclose(np.stack(conf), g["conf_" + tag], rtol=1e-5)
                    np.testing.assert_allclose(np.stack(coord), g["coord_" + tag], rtol=1e-5, atol=1e-6)
                    np.testing.assert_allclose(np.stack(size), g["size_" + tag], rtol=1e-5)
                    np.testing.assert_allclose(np.stack(feats), g["feat_" + tag], rtol=1e-5)


def test_predict_scene_differences(oracle):
    g = golden("find_match.npz")
    (s0, f0), (s1, f1) = _block_layers(oracle, g)
    for k in range(int(g["num_psd"])):
        use_feat = bool(g["psd%d_use_feat" % k])
        obj, g0, g1, _ = oracle.predict_scene_differences(
            s0, s1, f0 if use_feat else None, f1 if use_feat else None,
            set(g["psd%d_moved" % k].tolist()), g["psd%d_cands" % k].tolist(),
            confidence_threshold=0.0, contour_padding=0, contour_threshold=0.0, distance_threshold=0.05)
        assert (-1 if obj is None else obj) == int(g["psd%d_obj" % k]), k
        assert len(g0) == len(g["psd%d_g0" % k])
        if g0:
            np.testing.assert_allclose(np.stack(g0), g["psd%d_g0" % k], rtol=1e-5, atol=1e-6)
            np.testing.assert_allclose(np.stack(g1), g["psd%d_g1" % k], rtol=1e-5, atol=1e-6)
