"""Drop-in for the PROJECTION HALF of
/root/reference/mass/nn/applications/resnet_projection_layer.py: a 256-channel map at a
quarter of the camera resolution, fed the depth image sub-sampled [2::4, 2::4].

The ResNet-50 forward that produces the 56x56x256 feature image is out of scope
(SURVEY.md 8a9): pass the activations as observation["features"], or give the layer a
`feature_extractor` callable mapping observation["rgb"] to an [h, w, F] tensor.
"""
from typing import Any, Callable, Dict, Optional

import torch

from mass_b200.nn.base_projection_layer import BaseProjectionLayer


class ResNetProjectionLayer(BaseProjectionLayer):
    """Reference: resnet_projection_layer.py:62-141 (ctor), 159-211 (update)."""

    def __init__(self, camera_height: int = 224, camera_width: int = 224, vertical_fov: float = 90.0,
                 map_height: int = 256, map_width: int = 256, map_depth: int = 64,
                 feature_size: int = 256, dtype: torch.dtype = torch.float32, origin_y: float = 0.0,
                 origin_x: float = 0.0, origin_z: float = 0.0, grid_resolution: float = 0.05,
                 interpolation_weight: float = 0.5, initial_feature_map: torch.Tensor = None,
                 feature_extractor: Optional[Callable] = None, exact: bool = True):
        # the map sees a camera 4x coarser than the sensor (resnet_projection_layer.py:122-123)
        super().__init__(camera_height=camera_height // 4, camera_width=camera_width // 4,
                         vertical_fov=vertical_fov, map_height=map_height, map_width=map_width,
                         map_depth=map_depth, feature_size=feature_size, dtype=dtype, origin_y=origin_y,
                         origin_x=origin_x, origin_z=origin_z, grid_resolution=grid_resolution,
                         interpolation_weight=interpolation_weight,
                         initial_feature_map=initial_feature_map, exact=exact)
        self.feature_extractor = feature_extractor

    @staticmethod
    def subsample_depth(depth, feature_height):
        """depth[k//2::k, k//2::k] with k = depth rows // feature rows
        (resnet_projection_layer.py:203-210); accepts [..., H, W, 1]."""
        k = depth.shape[-3] // feature_height
        return depth[..., k // 2::k, k // 2::k, :]

    def update(self, observation: Dict[str, Any]):
        if "features" in observation:
            features = torch.as_tensor(observation["features"], dtype=torch.float32)
        elif self.feature_extractor is not None:
            features = self.feature_extractor(observation["rgb"])
        else:
            raise ValueError("ResNetProjectionLayer needs observation['features'] ([h, w, F] activations) "
                             "or a feature_extractor; the ResNet-50 forward is not part of this package")
        depth = torch.as_tensor(observation["depth"], dtype=torch.float32)
        return super().update(dict(position=observation["position"], yaw=observation["yaw"],
                                   elevation=observation["elevation"],
                                   depth=self.subsample_depth(depth, features.shape[0]), features=features))

    def update_batch(self, observations, fold=None):
        if isinstance(observations, (list, tuple)) and len(observations) == 0:
            return self
        if isinstance(observations, (list, tuple)):
            observations = {k: torch.stack([torch.as_tensor(o[k]) for o in observations])
                            for k in observations[0].keys()}
        features = torch.as_tensor(observations["features"], dtype=torch.float32)
        depth = torch.as_tensor(observations["depth"], dtype=torch.float32)
        return super().update_batch(dict(position=observations["position"], yaw=observations["yaw"],
                                         elevation=observations["elevation"],
                                         depth=self.subsample_depth(depth, features.shape[1]),
                                         features=features), fold=fold)

    def visualize(self, obs: Dict[str, Any], depth_slice: slice = slice(4, 32)):
        """The reference returns None here (resnet_projection_layer.py:245-269)."""
        return None
