"""Column summaries against the reference's own torch expressions evaluated on the CPU."""
import numpy as np
import pytest
import torch
from torch.nn import functional

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("F", [1, 5, 32, 54, 100])
def test_column_summary_matches_reference_expressions(F):
    from mass_b200.nn.base_projection_layer import BaseProjectionLayer
    from mass_b200.utils import navigation
    dev = torch.device("cuda:0")
    rng = np.random.default_rng(F)
    layer = BaseProjectionLayer(camera_height=8, camera_width=8, map_height=40, map_width=36, map_depth=24,
                                feature_size=F, grid_resolution=0.1).to(dev)
    data = (rng.standard_normal((40, 36, 24, F)) * (rng.random((40, 36, 24, 1)) < 0.05)).astype(np.float32)
    layer.data.copy_(torch.from_numpy(data))
    cpu = torch.from_numpy(data)
    amax, blocked = layer.column_summary()
    assert torch.equal(amax.cpu(), cpu.amax(dim=2))                              # agent.py:330-331
    assert torch.equal(blocked.cpu(), (torch.norm(cpu, p=1, dim=3) > 0.0).any(dim=2))
    for sl, pad in ((slice(4, 20), 3), (slice(0, 24), 0), (slice(10, 11), 1)):
        # mass/navigation_policy.py:207-221
        nav = torch.norm(cpu, p=1, dim=3) > 0.0
        nav = torch.logical_not(nav[:, :, sl].any(dim=2)).to(torch.float32)
        ref = 1 - functional.max_pool2d(1 - nav.unsqueeze(0), 2 * pad + 1, stride=1, padding=pad).squeeze(0)
        got = navigation.navigable_area(layer, padding=pad, depth_slice=sl)
        assert torch.equal(got.cpu(), ref)
    assert torch.equal(navigation.search_policy_input(layer).cpu(), cpu.amax(dim=2).unsqueeze(0).permute(0, 3, 1, 2))
    # a positive threshold: equal wherever the L1 norm is not within rounding of the threshold
    thr = 0.7
    l1 = torch.norm(cpu.double(), p=1, dim=3)
    _, blk = layer.column_summary(obstacle_threshold=thr)
    sure = ((l1 - thr).abs() > 1e-4).all(dim=2)
    assert torch.equal(blk.cpu()[sure], (l1 > thr).any(dim=2)[sure])
    with pytest.raises(ValueError):
        layer.column_summary(depth_slice=slice(0, 24, 2))


@pytest.mark.parametrize("F", [1, 3, 54, 70])
def test_top_down_matches_reference_expression(F):
    """BaseProjectionLayer.top_down against the reference's cumsum/arg-max/gather expression on the CPU (bit-exact)."""
    from mass_b200.nn.base_projection_layer import BaseProjectionLayer
    dev = torch.device("cuda:0")
    rng = np.random.default_rng(100 + F)
    layer = BaseProjectionLayer(camera_height=8, camera_width=8, map_height=33, map_width=29, map_depth=40,
                                feature_size=F, grid_resolution=0.1).to(dev)
    data = (rng.standard_normal((33, 29, 40, F)) * (rng.random((33, 29, 40, 1)) < 0.04)).astype(np.float32)
    data[0, 0] = 0.0                                  # an empty column
    data[1, 1, 39] = 1.0                              # filled at the very top
    layer.data.copy_(torch.from_numpy(data))
    cpu = torch.from_numpy(data)
    for sl in (slice(0, 32), slice(4, 32), None, slice(39, 40), slice(0, 1)):
        vol = cpu if sl is None else cpu[:, :, sl]
        mask = torch.ne(vol, 0).any(dim=-1, keepdim=True).to(vol.dtype)
        idx = (mask.cumsum(dim=-2) * mask).argmax(dim=-2, keepdim=True)
        ref = torch.gather(vol, -2, idx.expand(*vol.shape[:-2], 1, vol.shape[-1])).squeeze(-2)
        assert torch.equal(layer.top_down(depth_slice=sl).cpu(), ref), sl


def test_detections_to_ids_matches_reference_expression():
    """mass/thor/segmentation_config.py:314-334 evaluated with torch on the CPU vs the device kernel (bit-exact)."""
    from mass_b200.utils import perception
    dev = torch.device("cuda:0")
    rng = np.random.default_rng(77)
    H = W = 224
    for n in (0, 1, 37):
        masks = torch.from_numpy(rng.random((n, H, W)) < 0.15)
        yy, xx = np.mgrid[:H, :W]
        for i in range(n):                                          # blobs, so that overlaps and ties occur
            cy, cx, r = rng.integers(0, H), rng.integers(0, W), rng.integers(5, 60)
            masks[i] &= torch.from_numpy((yy - cy) ** 2 + (xx - cx) ** 2 < r * r)
        classes = torch.from_numpy(rng.integers(0, 54, n))
        scores = torch.from_numpy(rng.random(n).astype(np.float32))
        thr = 0.3
        seg = torch.zeros(H, W, 54)
        for i in range(n):
            if scores[i] < thr:
                continue
            seg[:, :, classes[i]] += masks[i].to(torch.float32)
        ref = seg.argmax(dim=2, keepdim=True)
        got = perception.detections_to_ids(masks.to(dev), classes.to(dev), scores.to(dev), thr)
        assert got.dtype == torch.int64 and tuple(got.shape) == (H, W, 1)
        assert torch.equal(got.cpu(), ref), n
