"""The hand-off from the detector to the semantic map.

Reference: /root/reference/mass/thor/segmentation_config.py:314-337 (SemanticRearrangeSensor.get_segmentation, the
Mask R-CNN branch): instance masks are accumulated per class into an [H, W, 54] float buffer, the arg-max class id per
pixel is taken and moved to the host as numpy, only for SemanticProjectionLayer.update to move it back and expand it
to one-hot.  Here the id image is produced on the device in one kernel and handed to the layer as `semantic`.
The detector itself (detectron2) is out of scope."""
import torch

from mass_b200 import _lib


def detections_to_ids(pred_masks, pred_classes, scores, detection_threshold: float, num_classes: int = 54):
    """pred_masks [n, H, W] (bool / uint8), pred_classes [n] integers, scores [n] floats, all on the GPU ->
    int64 id image [H, W, 1] on the GPU (the shape get_segmentation returns)."""
    device = _lib.require_cuda(pred_masks.device)
    if pred_masks.dim() != 3:
        raise ValueError("pred_masks must be [n, H, W], got %s" % (tuple(pred_masks.shape),))
    n, H, W = pred_masks.shape
    masks = pred_masks.to(torch.uint8).contiguous()
    classes = pred_classes.to(device=device, dtype=torch.int64).contiguous()
    scores = scores.to(device=device, dtype=torch.float32).contiguous()
    if classes.numel() != n or scores.numel() != n:
        raise ValueError("pred_classes and scores must have one entry per mask")
    ids = torch.empty(H, W, 1, dtype=torch.int64, device=device)
    _lib.check(_lib.lib().mb_masks_to_ids(_lib.stream_ptr(device), _lib.ptr(masks), _lib.ptr(classes), _lib.ptr(scores),
                                          n, H * W, int(num_classes), float(detection_threshold), _lib.ptr(ids)))
    return ids
