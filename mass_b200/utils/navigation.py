"""The two whole-map readers that sit next to the mapping path, with the reference's
signatures; the map sweep runs in one kernel (mb_column_summary), the tiny 2-D post-processing stays torch.

Reference: /root/reference/mass/navigation_policy.py:173-221 (NavigationPolicy.navigable_area) and
/root/reference/agent.py:330-331, 391-392 (input of the semantic search policy)."""
import torch
from torch.nn import functional


def navigable_area(feature_map, padding: int = 3, depth_slice: slice = None, obstacle_threshold: float = 0.0):
    """1 where the agent can stand, 0 where any voxel of the depth slice is occupied or within `padding` cells of
    one.  Same value as NavigationPolicy.navigable_area for the layer `feature_map` ([S0, S1] float)."""
    _, blocked = feature_map.column_summary(depth_slice=depth_slice, obstacle_threshold=obstacle_threshold,
                                            want_amax=False)
    navigable = torch.logical_not(blocked).to(dtype=feature_map.data.dtype)
    return 1 - functional.max_pool2d(1 - navigable.unsqueeze(0), 2 * padding + 1, stride=1, padding=padding).squeeze(0)


def search_policy_input(semantic_layer):
    """[1, F, S0, S1] = data.amax(dim=2).unsqueeze(0).permute(0, 3, 1, 2) (agent.py:330-331)."""
    amax, _ = semantic_layer.column_summary(want_blocked=False)
    return amax.unsqueeze(0).permute(0, 3, 1, 2)
