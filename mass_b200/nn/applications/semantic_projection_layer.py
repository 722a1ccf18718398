"""Drop-in for /root/reference/mass/nn/applications/semantic_projection_layer.py.

`update` takes the arg-max class-id image and fuses it as one-hot features WITHOUT
materialising the [H, W, 54] one-hot tensor: the voxel-reduce kernel reads the id
and synthesises the 0/1 channel values (same arithmetic, 54x less input traffic).
`find` extracts the instances of one class (see mass_b200/utils/instances.py).
"""
from typing import Any, Dict

import numpy as np
import torch

from mass_b200.nn.base_projection_layer import BaseProjectionLayer
from mass_b200.utils import instances


class SemanticProjectionLayer(BaseProjectionLayer):
    """Reference: semantic_projection_layer.py:9-362."""

    def __init__(self, camera_height: int = 224, camera_width: int = 224, vertical_fov: float = 90.0,
                 map_height: int = 256, map_width: int = 256, map_depth: int = 64, feature_size: int = 1,
                 dtype: torch.dtype = torch.float32, origin_y: float = 0.0, origin_x: float = 0.0,
                 origin_z: float = 0.0, grid_resolution: float = 0.05, interpolation_weight: float = 0.5,
                 initial_feature_map: torch.Tensor = None, class_to_colors=None, exact: bool = True):
        super().__init__(camera_height=camera_height, camera_width=camera_width, vertical_fov=vertical_fov,
                         map_height=map_height, map_width=map_width, map_depth=map_depth,
                         feature_size=feature_size, dtype=dtype, origin_y=origin_y, origin_x=origin_x,
                         origin_z=origin_z, grid_resolution=grid_resolution,
                         interpolation_weight=interpolation_weight,
                         initial_feature_map=initial_feature_map, exact=exact)
        # colour table for visualize(); the reference requires it (semantic_projection_layer.py:72,128)
        colors = torch.zeros(feature_size, 3) if class_to_colors is None else \
            torch.as_tensor(class_to_colors, dtype=torch.float32)
        self.register_buffer('class_to_colors', colors)
        self.boxes = None

    def reset(self, origin_y: float = 0.0, origin_x: float = 0.0, origin_z: float = 0.0):
        self.boxes = None
        super().reset(origin_y=origin_y, origin_x=origin_x, origin_z=origin_z)

    @staticmethod
    def _ids(semantic):
        ids = torch.as_tensor(semantic).to(torch.int64)
        return ids[..., 0] if ids.dim() >= 3 and ids.shape[-1] == 1 else ids

    def _validate(self, ids):
        """functional.one_hot raises on ids outside [0, feature_size) (semantic_projection_layer.py:203-214).  A host
        image (what the agent passes: numpy from the detector) is checked here, on the host.  A DEVICE
        image is not read back -- that would stall the stream on every frame; the kernel flags the bad id, adds
        nothing for that pixel, and `check()` raises the same error."""
        if not ids.is_cuda and ids.numel():
            lo, hi = torch.aminmax(ids)
            if int(lo) < 0 or int(hi) >= self.feature_size:
                raise RuntimeError("Class values must be in [0, %d)" % self.feature_size)

    def update(self, observation: Dict[str, Any]):
        """observation["semantic"]: [H, W, 1] integer class ids
        (semantic_projection_layer.py:203-214: one_hot(ids, feature_size).float())."""
        ids = self._ids(observation["semantic"])
        self._validate(ids)
        return super().update(dict(position=observation["position"], yaw=observation["yaw"],
                                   elevation=observation["elevation"], depth=observation["depth"],
                                   class_ids=ids))

    def update_batch(self, observations, fold=None):
        if isinstance(observations, (list, tuple)) and len(observations) == 0:
            return self
        if isinstance(observations, (list, tuple)):
            observations = {k: torch.stack([torch.as_tensor(o[k]) for o in observations])
                            for k in observations[0].keys()}
        ids = torch.as_tensor(observations["semantic"])
        if ids.is_cuda or ids.dtype not in (torch.int32, torch.int16, torch.uint8, torch.int8):
            ids = ids.to(torch.int64)           # (narrow host ids cross PCIe as they are and widen on the device)
        ids = ids.reshape(-1, self.camera_height, self.camera_width)
        # A batch of host images in the batched mode is not scanned on the host (a min/max pass over 500 frames of
        # int64 ids takes longer than copying them to the GPU): the kernels check every id they read anyway, the
        # host pipeline collects their error bits and raises before update_batch returns -- after the valid pixels
        # have been fused, where the reference raises before the offending frame.
        on_device_check = (not ids.is_cuda) and (not self.exact) and self.data.is_cuda and ids.shape[0] > 1 \
            and not torch.as_tensor(observations["depth"]).is_cuda
        if not on_device_check:
            self._validate(ids)
        return super().update_batch(dict(position=observations["position"], yaw=observations["yaw"],
                                         elevation=observations["elevation"], depth=observations["depth"],
                                         class_ids=ids), fold=fold, check_ids=on_device_check)

    def visualize(self, obs: Dict[str, Any], depth_slice: slice = slice(0, 32)):
        """Top-down arg-max class colours, white where empty, red boxes from the last find().
        Reference: semantic_projection_layer.py:218-255."""
        top = self.top_down(depth_slice=depth_slice)
        image = self.class_to_colors[top.argmax(dim=-1)]
        image = torch.where((top != 0).any(dim=-1, keepdim=True), image, torch.ones_like(image))
        image = image.cpu().numpy()
        for x, y, w, h in (self.boxes or []):
            x1, y1 = min(x + w, image.shape[1] - 1), min(y + h, image.shape[0] - 1)
            image[y, x:x1 + 1] = image[y1, x:x1 + 1] = np.array([1.0, 0.0, 0.0])
            image[y:y1 + 1, x] = image[y:y1 + 1, x1] = np.array([1.0, 0.0, 0.0])
        return image

    def find(self, semantic_category: int, confidence_threshold: float = 0.2, contour_padding: int = 3,
             contour_threshold: float = 0.0, feature_map=None):
        """Instances of one class: (confidences, coordinates, sizes, features-or-None), lists of
        tensors in OpenCV contour order.  Reference: semantic_projection_layer.py:257-362."""
        found = instances.find_instances(self, semantic_category, confidence_threshold, contour_padding,
                                         contour_threshold, feature_map)
        self.boxes = list(found.boxes)
        features = None if found.features is None else list(found.features)
        return list(found.confidences), list(found.coordinates), list(found.sizes), features
