"""Next to the mapping path, with the reference's semantics: the two whole-map readers
(one sweep, mb_column_summary), the obstacle padding, and the navigation graph's construction / refresh, whose
traversability tests run as kernels over the whole node lattice (reset) or over every node and edge of an existing
graph (update) instead of a Python loop of tensor slices per edge.

Reference: /root/reference/mass/navigation_policy.py:173-221 (navigable_area), :223-285 (reset_navigation_graph),
:287-341 (update_navigation_graph) and /root/reference/agent.py:330-331, 391-392 (input of the semantic search policy).
The graph itself stays a networkx.Graph with the reference's node names (x, y): that is what its path planner reads."""
import numpy as np
import torch

from mass_b200 import _lib


def navigable_area(feature_map, padding: int = 3, depth_slice: slice = None, obstacle_threshold: float = 0.0):
    """1 where the agent can stand, 0 where any voxel of the depth slice is occupied or within `padding` cells of
    one.  Same value as NavigationPolicy.navigable_area for the layer `feature_map` ([S0, S1] float)."""
    _, blocked = feature_map.column_summary(depth_slice=depth_slice, obstacle_threshold=obstacle_threshold,
                                            want_amax=False)
    device = _lib.require_cuda(blocked.device)
    blocked = blocked.to(torch.uint8).contiguous()
    S0, S1 = blocked.shape
    out = torch.empty(S0, S1, dtype=feature_map.data.dtype, device=device)
    _lib.check(_lib.lib().mb_navigable_area(_lib.stream_ptr(device), _lib.ptr(blocked), S0, S1, int(padding), _lib.ptr(out)))
    return out


def search_policy_input(semantic_layer):
    """[1, F, S0, S1] = data.amax(dim=2).unsqueeze(0).permute(0, 3, 1, 2) (agent.py:330-331)."""
    amax, _ = semantic_layer.column_summary(want_blocked=False)
    return amax.unsqueeze(0).permute(0, 3, 1, 2)


def lattice_offset(feature_map, step_size):
    """(offset_x, offset_y): the lattice is shifted so that the voxel at the map's world origin is a node
    (navigation_policy.py:262-268)."""
    origin = torch.tensor([[feature_map.origin_x, feature_map.origin_y]], dtype=torch.float32)
    cell = feature_map.world_to_map(origin)                          # clamping never moves the origin of a map
    return int(cell[0, 0]) % step_size, int(cell[0, 1]) % step_size


def reset_navigation_graph(feature_map, step_size: int = 5, padding: int = 3, depth_slice: slice = None,
                           obstacle_threshold: float = 0.0):
    """A fresh graph over the node lattice: an edge wherever the cells between two neighbouring nodes are all
    navigable.  Returns (networkx.Graph, navigable_area)."""
    import networkx
    nav = navigable_area(feature_map, padding=padding, depth_slice=depth_slice, obstacle_threshold=obstacle_threshold)
    device = nav.device
    S0, S1 = nav.shape
    off_x, off_y = lattice_offset(feature_map, step_size)
    ny, nx = -(-(S0 - off_y) // step_size), -(-(S1 - off_x) // step_size)
    node_ok = torch.empty(ny, nx, dtype=torch.uint8, device=device)
    edge_ok = torch.empty(ny, nx, 2, dtype=torch.uint8, device=device)
    _lib.check(_lib.lib().mb_nav_graph_lattice(_lib.stream_ptr(device), _lib.ptr(nav.contiguous()), S0, S1, off_y, off_x,
                                               int(step_size), _lib.ptr(node_ok), _lib.ptr(edge_ok)))
    edges = edge_ok.cpu().numpy().astype(bool)
    graph = networkx.Graph()
    ii = off_y + step_size * np.arange(ny)
    jj = off_x + step_size * np.arange(nx)
    # insertion order of the reference's double loop: rows outer, columns inner, down before right
    for a, b, d in zip(*np.nonzero(edges)):
        i, j = int(ii[a]), int(jj[b])
        graph.add_edge((j, i), (j, i + step_size) if d == 0 else (j + step_size, i))
    return graph, nav


def update_navigation_graph(graph, feature_map, padding: int = 3, depth_slice: slice = None,
                            obstacle_threshold: float = 0.0):
    """Drops the nodes that are now obstructed and the edges with an obstructed cell between their ends; everything else
    stays (isolated nodes included), exactly as the reference's in-place refresh.  Returns the navigable area."""
    nav = navigable_area(feature_map, padding=padding, depth_slice=depth_slice, obstacle_threshold=obstacle_threshold)
    device = nav.device
    S0, S1 = nav.shape
    nodes = list(graph.nodes())
    edges = list(graph.edges())
    rects = [(i, i, j, j) for (j, i) in nodes] + \
            [(min(i, y), max(i, y), min(j, x), max(j, x)) for (j, i), (x, y) in edges]
    if not rects:
        return nav
    r = torch.tensor(rects, dtype=torch.int32).to(device)
    clear = torch.empty(len(rects), dtype=torch.uint8, device=device)
    _lib.check(_lib.lib().mb_nav_rects_clear(_lib.stream_ptr(device), _lib.ptr(nav.contiguous()), S0, S1, _lib.ptr(r),
                                             len(rects), _lib.ptr(clear)))
    clear = clear.cpu().numpy().astype(bool)
    # a node is dropped when its cell is 0 (navigable areas are 0 / 1 images, so "not 1" is "0")
    graph.remove_nodes_from([n for n, ok in zip(nodes, clear[:len(nodes)]) if not ok])
    graph.remove_edges_from([e for e, ok in zip(edges, clear[len(nodes):]) if not ok and graph.has_edge(*e)])
    return nav
