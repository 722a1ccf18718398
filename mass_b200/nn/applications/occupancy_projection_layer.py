"""Drop-in for /root/reference/mass/nn/applications/occupancy_projection_layer.py:
a one-channel map whose feature image is all ones (occupied where depth lands)."""
from typing import Any, Dict

import numpy as np
import torch

from mass_b200.nn.base_projection_layer import BaseProjectionLayer


class OccupancyProjectionLayer(BaseProjectionLayer):
    """Reference: occupancy_projection_layer.py:9-189 (feature_size is 1 there too)."""

    def update(self, observation: Dict[str, Any]):
        """features = ones_like(depth)  (occupancy_projection_layer.py:158-161)."""
        depth = torch.as_tensor(observation["depth"], dtype=torch.float32)
        return super().update(dict(position=observation["position"], yaw=observation["yaw"],
                                   elevation=observation["elevation"], depth=depth,
                                   features=torch.ones_like(depth).reshape(
                                       self.camera_height, self.camera_width, 1)))

    def update_batch(self, observations, fold=None):
        if isinstance(observations, (list, tuple)) and len(observations) == 0:
            return self
        if isinstance(observations, (list, tuple)):
            observations = {k: torch.stack([torch.as_tensor(o[k]) for o in observations])
                            for k in ("position", "yaw", "elevation", "depth")}
        depth = torch.as_tensor(observations["depth"], dtype=torch.float32)
        T = depth.numel() // (self.camera_height * self.camera_width)
        return super().update_batch(dict(position=observations["position"], yaw=observations["yaw"],
                                         elevation=observations["elevation"], depth=depth,
                                         features=torch.ones_like(depth).reshape(
                                             T, self.camera_height, self.camera_width, 1)), fold=fold)

    def visualize(self, obs: Dict[str, Any], depth_slice: slice = slice(4, 32)):
        """Occupied columns in black, the agent's cell in red.  The reference draws the
        planned path with OpenCV (occupancy_projection_layer.py:165-189 ->
        mass/utils/visualization.py, debug rendering, out of scope); this keeps the
        return type (an [S0, S1, 3] float image) without that dependency."""
        vol = self.data if depth_slice is None else self.data[:, :, depth_slice]
        occupied = (vol != 0).any(dim=-1).any(dim=-1, keepdim=True).to(torch.float32).cpu().numpy()
        image = 1.0 - np.tile(occupied, (1, 1, 3))
        if obs is not None and "position" in obs:
            cell = self.world_to_map(torch.as_tensor(obs["position"], dtype=torch.float32)[:2])
            x, y = int(cell[0]), int(cell[1])
            image[max(y - 1, 0):y + 2, max(x - 1, 0):x + 2] = np.array([1.0, 0.0, 0.0])
        return image
