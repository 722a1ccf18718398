"""Coordinate transforms and the navigation graph: the oracle restatement against golden vectors
from the UNMODIFIED reference (CPU), and the kernels against the same vectors (GPU).
Reference: mass/nn/base_projection_layer.py:381-547, mass/navigation_policy.py:173-341."""
import numpy as np
import pytest
import torch

from conftest import golden, golden_kwargs


class _Shell:
    pass


def _oracle_layer(oracle, g, data=None):
    kw = golden_kwargs(g)
    L = oracle.OracleLayer(**kw)
    if data is not None:
        L.data[...] = data
    return L


def _args(g, tag):
    pad, s0, s1, step = (int(v) for v in g["args_" + tag])
    return pad, (None if s0 < 0 else slice(s0, s1)), float(g["thr_" + tag]), step


def test_oracle_coordinate_transforms_match_the_reference(oracle):
    g = golden("navigation.npz")
    L = _oracle_layer(oracle, g)
    assert np.array_equal(oracle.world_to_map(L, g["world"]), g["cells3"])
    assert np.array_equal(oracle.world_to_map(L, g["world"][:, :2]), g["cells2"])
    assert np.array_equal(oracle.map_to_world(L, g["mapc"]), g["back3"])


@pytest.mark.parametrize("tag", ["a", "b"])
def test_oracle_navigation_graph_matches_the_reference(oracle, tag):
    g = golden("navigation.npz")
    pad, sl, thr, step = _args(g, tag)
    L = _oracle_layer(oracle, g, g["data"])
    nav = oracle.navigable_area(L, pad, sl, thr)
    assert np.array_equal(nav, g["nav_" + tag])
    import networkx
    edges = oracle.navigation_graph_edges(L, nav, step)
    graph = networkx.Graph()
    graph.add_edges_from(edges)                        # insertion order decides the order networkx iterates in
    assert [list(a) + list(b) for a, b in graph.edges()] == g["edges_" + tag].tolist()
    assert [list(n) for n in graph.nodes()] == g["nodes_" + tag].tolist()
    L.data[...] = g["data2"]
    nav2 = oracle.navigable_area(L, pad, sl, thr)
    nodes, kept = oracle.update_navigation_graph([tuple(n) for n in g["nodes_" + tag].tolist()], edges, nav2)
    assert [list(n) for n in nodes] == g["nodes2_" + tag].tolist()
    norm = lambda e: tuple(sorted((tuple(e[:2]), tuple(e[2:]))))                        # noqa: E731  (u, v) == (v, u)
    assert sorted(norm(list(a) + list(b)) for a, b in kept) == sorted(norm(e) for e in g["edges2_" + tag].tolist())


@pytest.mark.gpu
def test_gpu_coordinate_transforms_bit_exact():
    from mass_b200.nn.base_projection_layer import BaseProjectionLayer
    g = golden("navigation.npz")
    dev = torch.device("cuda:0")
    L = BaseProjectionLayer(**golden_kwargs(g)).to(dev)
    assert np.array_equal(L.world_to_map(torch.from_numpy(g["world"])).cpu().numpy(), g["cells3"])
    assert np.array_equal(L.world_to_map(torch.from_numpy(g["world"][:, :2]).to(dev)).cpu().numpy(), g["cells2"])
    assert np.array_equal(L.map_to_world(torch.from_numpy(g["mapc"])).cpu().numpy(), g["back3"])
    # any leading shape, and integer cells as the callers pass them (semantic find, navigation policy)
    cells = torch.from_numpy(g["cells3"]).reshape(10, 30, 3).to(dev)
    got = L.map_to_world(cells)
    ref = BaseProjectionLayer(**golden_kwargs(g)).map_to_world(torch.from_numpy(g["cells3"]).reshape(10, 30, 3))
    assert tuple(got.shape) == (10, 30, 3) and torch.equal(got.cpu(), ref)
    big = torch.rand(200000, 3, device=dev) * 8 - 4
    assert torch.equal(L.world_to_map(big).cpu(), BaseProjectionLayer(**golden_kwargs(g)).world_to_map(big.cpu()))
    assert L.world_to_map(torch.zeros(0, 3)).shape == (0, 3)


@pytest.mark.gpu
@pytest.mark.parametrize("tag", ["a", "b"])
def test_gpu_navigation_graph_matches_the_reference(tag):
    from mass_b200.nn.base_projection_layer import BaseProjectionLayer
    from mass_b200.utils import navigation
    g = golden("navigation.npz")
    pad, sl, thr, step = _args(g, tag)
    dev = torch.device("cuda:0")
    L = BaseProjectionLayer(**golden_kwargs(g)).to(dev)
    L.data.copy_(torch.from_numpy(g["data"]))
    graph, nav = navigation.reset_navigation_graph(L, step_size=step, padding=pad, depth_slice=sl, obstacle_threshold=thr)
    assert np.array_equal(nav.cpu().numpy(), g["nav_" + tag])
    assert [list(a) + list(b) for a, b in graph.edges()] == g["edges_" + tag].tolist()
    assert [list(n) for n in graph.nodes()] == g["nodes_" + tag].tolist()
    L.data.copy_(torch.from_numpy(g["data2"]))
    L.mark_dirty()
    navigation.update_navigation_graph(graph, L, padding=pad, depth_slice=sl, obstacle_threshold=thr)
    assert [list(n) for n in graph.nodes()] == g["nodes2_" + tag].tolist()
    assert [list(a) + list(b) for a, b in graph.edges()] == g["edges2_" + tag].tolist()
