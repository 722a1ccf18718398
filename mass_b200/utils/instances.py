"""Instance extraction for SemanticProjectionLayer.find and the matching primitives of
predict_scene_differences, on the kernels of libmassb200 (no CPU path for the tensor work).

Reference: /root/reference/mass/nn/applications/semantic_projection_layer.py:257-362 (find),
/root/reference/mass/utils/experimentation.py:261-287 (cost matrices + assignment).

Contour extraction stays on the host with OpenCV exactly as in the reference
(cv2.findContours / cv2.boundingRect, semantic_projection_layer.py:323-328): OpenCV decides the
number and ORDER of the instances, so using it keeps instance indices identical by construction.
"""
import ctypes
from collections import namedtuple

import numpy as np
import torch

from mass_b200 import _lib

Instances = namedtuple("Instances", "boxes confidences coordinates sizes features")

_ws = _lib.Workspace()


def class_presence(layer, semantic_category, contour_padding, contour_threshold):
    """uint8 [S0, S1] CUDA image: any over z of (box mean of data[..., c]) > threshold
    (semantic_projection_layer.py:309-317)."""
    data = layer.data
    device = _lib.require_cuda(data.device)
    S0, S1, S2, F = data.shape
    if not 0 <= int(semantic_category) < F:
        raise IndexError("semantic_category %d is outside [0, %d)" % (semantic_category, F))
    L = _lib.lib()
    image = torch.empty(S0, S1, dtype=torch.uint8, device=device)
    ws = _ws.get(L.mb_class_presence_workspace_bytes(S0, S1, S2, int(contour_padding)), device)
    _lib.check(L.mb_class_presence(_lib.stream_ptr(device), _lib.ptr(data), S0, S1, S2, F, int(semantic_category),
                                   int(contour_padding), float(contour_threshold), _lib.ptr(image), _lib.ptr(ws),
                                   ws.numel()))
    return image


def contour_boxes(threshold_image):
    """Host step of find(): OpenCV contours -> bounding boxes (x, y, w, h), OpenCV's order."""
    import cv2
    contours = cv2.findContours(np.ascontiguousarray(threshold_image), cv2.RETR_LIST, cv2.CHAIN_APPROX_SIMPLE)[0]
    return [tuple(int(v) for v in cv2.boundingRect(c)) for c in contours]


def pool_boxes(layer, semantic_category, boxes, feature_map=None):
    """[n, 5 + FF] CUDA rows {confidence, x, y, z, size, feature...} for bounding boxes over the
    full map depth (semantic_projection_layer.py:329-357)."""
    data = layer.data
    device = _lib.require_cuda(data.device)
    S0, S1, S2, F = data.shape
    feat, FF = None, 0
    if feature_map is not None:
        feat = feature_map.data
        if not feat.is_cuda:
            raise RuntimeError("the instance feature map must live on the GPU (it is %s); the reference keeps "
                               "it on the host only because 13.5 GiB did not fit its GPU" % feat.device)
        if feat.device != device or tuple(feat.shape[:3]) != (S0, S1, S2) or feat.dtype != torch.float32:
            raise ValueError("feature map must be float32 [%d, %d, %d, FF] on %s" % (S0, S1, S2, device))
        FF = int(feat.shape[3])
    n = len(boxes)
    out = torch.empty(n, 5 + FF, dtype=torch.float32, device=device)
    if n == 0:
        return out
    mx, my, mz = layer.cell_centres()
    boxes_d = torch.tensor(boxes, dtype=torch.int32).reshape(n, 4).to(device)
    _lib.check(_lib.lib().mb_instance_pool(
        _lib.stream_ptr(device), _lib.ptr(boxes_d), n, _lib.ptr(data), S0, S1, S2, F, int(semantic_category),
        _lib.ptr(feat), FF, _lib.ptr(mx.contiguous()), _lib.ptr(my.contiguous()), _lib.ptr(mz.contiguous()),
        _lib.ptr(out)))
    return out


def _presence_image(layer, semantic_category, contour_padding, contour_threshold):
    """Host uint8 [S0, S1] image of one class.  Without smoothing (the agent's default, agent.py:846-847)
    any_z(v > thr) == (max_z v > thr), so ONE sweep over the map (mb_column_summary) serves all classes; the
    per-class images are kept on the host until the map changes."""
    if contour_padding != 0:
        return class_presence(layer, semantic_category, contour_padding, contour_threshold).cpu().numpy()
    state = (layer.map_state(), float(contour_threshold))
    cache = getattr(layer, "_presence_cache", None)
    if cache is None or cache[0] != state:
        amax, _ = layer.column_summary(want_blocked=False)
        stack = (amax > contour_threshold).to(torch.uint8).permute(2, 0, 1).contiguous().cpu().numpy()
        cache = (state, stack)
        layer._presence_cache = cache
    if not 0 <= int(semantic_category) < cache[1].shape[0]:
        raise IndexError("semantic_category %d is outside [0, %d)" % (semantic_category, cache[1].shape[0]))
    return cache[1][int(semantic_category)]


def _pool_rows(layer, boxes5, feature_map):
    """[n, 5 + FF] rows for boxes (x, y, w, h, class) of any mix of classes, one launch."""
    data = layer.data
    device = _lib.require_cuda(data.device)
    S0, S1, S2, F = data.shape
    feat, FF = None, 0
    if feature_map is not None:
        feat = feature_map.data
        if not feat.is_cuda:
            raise RuntimeError("the instance feature map must live on the GPU (it is %s); the reference keeps "
                               "it on the host only because 13.5 GiB did not fit its GPU" % feat.device)
        if feat.device != device or tuple(feat.shape[:3]) != (S0, S1, S2) or feat.dtype != torch.float32:
            raise ValueError("feature map must be float32 [%d, %d, %d, FF] on %s" % (S0, S1, S2, device))
        FF = int(feat.shape[3])
    n = len(boxes5)
    out = torch.empty(n, 5 + FF, dtype=torch.float32, device=device)
    if n == 0:
        return out
    mx, my, mz = layer.cell_centres()
    boxes_d = torch.tensor(boxes5, dtype=torch.int32).reshape(n, 5).to(device)
    _lib.check(_lib.lib().mb_instance_pool(
        _lib.stream_ptr(device), _lib.ptr(boxes_d), n, _lib.ptr(data), S0, S1, S2, F, -1,
        _lib.ptr(feat), FF, _lib.ptr(mx.contiguous()), _lib.ptr(my.contiguous()), _lib.ptr(mz.contiguous()),
        _lib.ptr(out)))
    return out


def _instances_from_rows(boxes, rows, keep, with_features):
    kept = [i for i in range(len(boxes)) if keep[i]]
    return Instances(
        boxes=[boxes[i] for i in kept],
        confidences=[rows[i, 0] for i in kept],
        coordinates=[rows[i, 1:4] for i in kept],
        sizes=[rows[i, 4] for i in kept],
        features=[rows[i, 5:] for i in kept] if with_features else None)


def find_instances(layer, semantic_category, confidence_threshold, contour_padding, contour_threshold,
                   feature_map):
    """find() is a pure function of the two maps and its arguments: results are cached until a map changes
    (the agent's matching loop asks for the same classes again after every rearranged object).  Without
    smoothing (the agent's default) the first query after a map change extracts the instances of ALL classes:
    one map sweep for the presence images, OpenCV contours per class on the host, one pooling launch and one
    device-to-host copy for every box of every class."""
    c = int(semantic_category)
    params = (float(confidence_threshold), int(contour_padding), float(contour_threshold))
    maps = (layer.map_state(), None if feature_map is None else (id(feature_map), feature_map.map_state()))
    memo = getattr(layer, "_find_cache", None)
    if memo is None or memo[0] != maps[0]:
        memo = (maps[0], {})
        layer._find_cache = memo
    key = (c,) + params + maps
    if key in memo[1]:
        return memo[1][key]
    if contour_padding == 0:
        num_classes = layer.data.shape[3]
        if not 0 <= c < num_classes:
            raise IndexError("semantic_category %d is outside [0, %d)" % (c, num_classes))
        # contour boxes of EVERY class, once per state of the semantic map (one sweep for the presence images, OpenCV
        # on the host per class): kept, so that a class served later costs one pooling launch and no contour pass
        boxes_key = ("boxes", float(contour_threshold))
        all_boxes = memo[1].get(boxes_key)
        if all_boxes is None:
            all_boxes = {k: contour_boxes(_presence_image(layer, k, 0, contour_threshold)) for k in range(num_classes)}
            memo[1][boxes_key] = all_boxes
        bulk_key = ("bulk",) + params + maps              # the shared pooling launch, once per (arguments, maps)
        first = bulk_key not in memo[1]
        memo[1][bulk_key] = True
        # classes whose boxes cover a large part of the map (the background class spans the whole room) are
        # left out of the shared launch and pooled on demand: one CTA walks a box
        wanted = [k for k in range(num_classes)
                  if k == c or (first and sum(b[2] * b[3] for b in all_boxes[k]) <= 4096)]
        wanted = [k for k in wanted if (k,) + params + maps not in memo[1]]
        boxes5 = [b + (k,) for k in wanted for b in all_boxes[k]]
        rows = _pool_rows(layer, boxes5, feature_map)
        keep = (rows[:, 0] > confidence_threshold).cpu().numpy() if boxes5 else np.zeros(0, bool)
        start = 0
        for k in wanted:
            bk = all_boxes[k]
            sl = slice(start, start + len(bk))
            memo[1][(k,) + params + maps] = _instances_from_rows(bk, rows[sl], keep[sl], feature_map is not None)
            start += len(bk)
        return memo[1][key]
    image = _presence_image(layer, c, contour_padding, contour_threshold)
    boxes = contour_boxes(image)
    rows = pool_boxes(layer, c, boxes, feature_map)
    keep = (rows[:, 0] > confidence_threshold).cpu().numpy() if len(boxes) else np.zeros(0, bool)
    memo[1][key] = found = _instances_from_rows(boxes, rows, keep, feature_map is not None)
    return found


def pairwise_l2(a, b):
    """[n, m] distances ||a_i - b_j||_2 from direct differences
    (torch.linalg.norm(a.unsqueeze(1) - b.unsqueeze(0), dim=2), experimentation.py:261-265)."""
    device = _lib.require_cuda(a.device)
    a = a.to(torch.float32).contiguous()
    b = b.to(device=device, dtype=torch.float32).contiguous()
    if a.dim() != 2 or b.dim() != 2 or a.shape[1] != b.shape[1]:
        raise ValueError("pairwise_l2 needs [n, d] and [m, d], got %s and %s" % (tuple(a.shape), tuple(b.shape)))
    out = torch.empty(a.shape[0], b.shape[0], dtype=torch.float32, device=device)
    _lib.check(_lib.lib().mb_pairwise_l2(_lib.stream_ptr(device), _lib.ptr(a), a.shape[0], _lib.ptr(b), b.shape[0],
                                         a.shape[1], _lib.ptr(out)))
    return out


def cosine_best_match(a, b, tensor_cores=None):
    """(best [n] int64, similarity [n] float32): for every row of a the row of b with the largest cosine
    similarity, first maximum on ties.  Large matrices (>= 1024 rows on both sides; `tensor_cores` forces either
    path) go through the tcgen05 kernel: 3 x TF32 ranks the candidates, float64 decides, so both paths return the same
    indices.  An additional op: the reference matches by L2 distance + assignment
    (predict_scene_differences), which is what `pairwise_l2` + `linear_sum_assignment` reproduce."""
    device = _lib.require_cuda(a.device)
    a = a.to(torch.float32).contiguous()
    b = b.to(device=device, dtype=torch.float32).contiguous()
    if a.dim() != 2 or b.dim() != 2 or a.shape[1] != b.shape[1]:
        raise ValueError("cosine_best_match needs [n, d] and [m, d], got %s and %s" % (tuple(a.shape), tuple(b.shape)))
    best = torch.empty(a.shape[0], dtype=torch.int64, device=device)
    sim = torch.empty(a.shape[0], dtype=torch.float32, device=device)
    if tensor_cores is None:
        # a real dense contraction only from a few thousand instances on; below that the job is latency-sized
        tensor_cores = a.shape[0] >= 1024 and b.shape[0] >= 1024 and a.shape[1] >= 32
    if tensor_cores and a.shape[0] > 0:
        L = _lib.lib()
        nbytes = int(L.mb_cosine_best_match_tc_workspace_bytes(a.shape[0], b.shape[0], a.shape[1]))
        ws = _ws.get(nbytes, device)
        _lib.check(L.mb_cosine_best_match_tc(_lib.stream_ptr(device), _lib.ptr(a), a.shape[0], _lib.ptr(b), b.shape[0],
                                             a.shape[1], _lib.ptr(best), _lib.ptr(sim), _lib.ptr(ws), nbytes))
        return best, sim
    _lib.check(_lib.lib().mb_cosine_best_match(_lib.stream_ptr(device), _lib.ptr(a), a.shape[0], _lib.ptr(b),
                                               b.shape[0], a.shape[1], _lib.ptr(best), _lib.ptr(sim)))
    return best, sim


def linear_sum_assignment(cost):
    """scipy.optimize.linear_sum_assignment on a CUDA cost matrix (float32 or float64): returns
    (rows, cols) int64 numpy arrays, rows ascending, same tie behaviour as scipy (one CTA runs the
    shortest-augmenting-path solver with float64 duals)."""
    device = _lib.require_cuda(cost.device)
    if cost.dim() != 2:
        raise ValueError("expected a matrix (2-D array), got a %d array" % cost.dim())
    if cost.dtype not in (torch.float32, torch.float64):
        cost = cost.to(torch.float64)
    cost = cost.contiguous()
    n, m = cost.shape
    k = min(n, m)
    if k == 0:
        return np.zeros(0, np.int64), np.zeros(0, np.int64)
    if bool(torch.isnan(cost).any() | (cost == -float("inf")).any()):     # scipy rejects NaN and -inf
        raise ValueError("matrix contains invalid numeric entries")
    L = _lib.lib()
    pairs = torch.empty(2, k, dtype=torch.int64, device=device)
    status = torch.zeros(1, dtype=torch.int32, device=device)
    ws = _ws.get(L.mb_lsap_workspace_bytes(n, m), device)
    c32 = cost if cost.dtype == torch.float32 else None
    c64 = cost if cost.dtype == torch.float64 else None
    _lib.check(L.mb_lsap(_lib.stream_ptr(device), _lib.ptr(c32), _lib.ptr(c64), n, m, _lib.ptr(pairs[0]),
                         _lib.ptr(pairs[1]), _lib.ptr(status), _lib.ptr(ws), ws.numel()))
    host = pairs.cpu().numpy()
    if int(status.item()) != 0:
        raise ValueError("cost matrix is infeasible")
    return host[0].copy(), host[1].copy()
