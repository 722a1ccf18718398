"""Drop-in for the matcher of /root/reference/mass/utils/experimentation.py:169-313
(predict_scene_differences) and for the rearrangement ordering of /root/reference/agent.py:455-465.

Same arguments, control flow and return values as the reference; the instance extraction,
both cost matrices and the assignment run on the kernels of libmassb200.  The simulator-facing
rest of the reference module (restart loop, Unity time-outs, ground-truth differences) is out of
scope.
"""
from typing import Iterable, Set

import torch

from mass_b200.utils import instances

# class tables of /root/reference/mass/thor/segmentation_config.py:43-117: id 0 is
# "OccupiedSpace", ids 1..43 are the pickable classes, ids 44..53 the openable ones
NUM_CLASSES = 54
ID_TO_PICKABLE = [1 <= i <= 43 for i in range(NUM_CLASSES)]
ID_TO_OPENABLE = [44 <= i <= 53 for i in range(NUM_CLASSES)]


def match_instances(feature0, feature1, goal0, goal1, size0, size1, object_pickable):
    """Cost matrices + assignment of one class (experimentation.py:261-287).
    Returns (rows, cols, distance) with rows/cols int64 numpy arrays and distance a CUDA matrix."""
    if feature0 is not None and feature1 is not None:
        deformation = instances.pairwise_l2(torch.stack(feature0, dim=0), torch.stack(feature1, dim=0))
    else:
        size0, size1 = torch.stack(size0, dim=0), torch.stack(size1, dim=0)
        deformation = (size0.unsqueeze(1) - size1.unsqueeze(0)).abs()
    distance = instances.pairwise_l2(goal0, goal1)
    rows, cols = instances.linear_sum_assignment(deformation if object_pickable else distance)
    return rows, cols, distance


def predict_scene_differences(semantic_projection_layer0, semantic_projection_layer1,
                              resnet_projection_layer0, resnet_projection_layer1,
                              objects_moved: Set[int], object_ids_to_move_pred: Iterable[int],
                              confidence_threshold: float = 0.2, contour_padding: int = 3,
                              contour_threshold: float = 0.0, distance_threshold: float = 0.0,
                              deformation_threshold: float = 0.0):
    """Which object class differs between two semantic maps, and where its instances are.
    Returns (object_to_move or None, goals in map 0, goals in map 1) -- see the reference
    docstring (experimentation.py:180-229); deformation_threshold is unused there as well."""
    object_to_move = None
    object_goals0, object_goals1 = [], []
    for candidate_object in object_ids_to_move_pred:
        object_pickable = ID_TO_PICKABLE[candidate_object]
        object_openable = ID_TO_OPENABLE[candidate_object]
        if candidate_object in objects_moved or not (object_pickable or object_openable):
            continue
        kw = dict(contour_padding=contour_padding, contour_threshold=contour_threshold,
                  confidence_threshold=confidence_threshold)
        conf0, goal0, size0, feature0 = semantic_projection_layer0.find(
            candidate_object, feature_map=resnet_projection_layer0, **kw)
        conf1, goal1, size1, feature1 = semantic_projection_layer1.find(
            candidate_object, feature_map=resnet_projection_layer1, **kw)
        if len(conf0) == 0 or len(conf1) == 0:
            continue
        # the assignment of a class is a pure function of the four maps: the agent's loop (agent.py:424-450) asks
        # again after every rearranged object, so it is kept -- with the stacked goal tensors -- until a map changes
        state = tuple(None if m is None else (id(m), m.map_state()) for m in
                      (semantic_projection_layer0, semantic_projection_layer1, resnet_projection_layer0,
                       resnet_projection_layer1))
        memo = getattr(semantic_projection_layer0, "_match_cache", None)
        if memo is None or memo[0] != state:
            memo = (state, {})
            semantic_projection_layer0._match_cache = memo
        key = (candidate_object, float(confidence_threshold), int(contour_padding), float(contour_threshold),
               float(distance_threshold))
        if key not in memo[1]:
            goal0, goal1 = torch.stack(goal0, dim=0), torch.stack(goal1, dim=0)
            rows, cols, distance = match_instances(feature0, feature1, goal0, goal1, size0, size1, object_pickable)
            memo[1][key] = (rows, cols, (distance > distance_threshold).cpu().numpy(), goal0, goal1)
        rows, cols, far, goal0, goal1 = memo[1][key]
        for instance0, instance1 in zip(rows, cols):
            if (object_pickable and far[instance0, instance1]) or object_openable:
                object_to_move = candidate_object
                object_goals0.append(goal0[instance0])
                object_goals1.append(goal1[instance1])
        if object_to_move is not None:
            break
    return object_to_move, object_goals0, object_goals1


def order_goals(object_goals0, object_goals1):
    """agent.py:455-465: visit first the pair whose nearest counterpart is farthest away.
    Returns the permutation (int64 numpy) applied to both goal lists."""
    goal0, goal1 = torch.stack(object_goals0, dim=0), torch.stack(object_goals1, dim=0)
    distance = instances.pairwise_l2(goal0, goal1)
    return distance.amin(dim=1).argsort(descending=True).cpu().numpy()
