"""The plugin contract of the mapping path: nine abstract methods every
projection layer implements.  Mirrors /root/reference/mass/nn/projection_layer.py:4-256
(same method names and meaning) so that agent.py / NavigationPolicy can take a
mass_b200 layer wherever they take a reference one."""
import abc


class ProjectionLayer(abc.ABC):
    """A voxel feature map fed by posed RGB-D observations.

    Conventions shared by all implementations (reference lines in brackets):
    world frame is (x, y, z-up); the map tensor is stored [y flipped, x, z, F];
    map coordinates passed to the coordinate helpers are in (x, y, z) order.
    """

    @abc.abstractmethod
    def get_feature_map(self, *args, **kwargs):
        """The dense voxel tensor [map_height, map_width, map_depth, F]. [56-70]"""

    @abc.abstractmethod
    def update(self, *args, **kwargs):
        """Fuse one observation dict (position, yaw, elevation, depth, features)
        into the map; returns self. [72-99]"""

    @abc.abstractmethod
    def reset(self, *args, **kwargs):
        """Zero the map and move its origin. [101-121]"""

    @abc.abstractmethod
    def top_down(self, *args, **kwargs):
        """Feature image of the top-most occupied voxel per (y, x) cell. [123-143]"""

    @abc.abstractmethod
    def clamp_to_world(self, *args, **kwargs):
        """Clamp world coordinates to the span of voxel centres. [145-165]"""

    @abc.abstractmethod
    def clamp_to_map(self, *args, **kwargs):
        """Clamp map coordinates to [0, size - 1]. [167-187]"""

    @abc.abstractmethod
    def map_to_world(self, *args, **kwargs):
        """Map (x, y, z) cell coordinates -> world coordinates. [189-209]"""

    @abc.abstractmethod
    def world_to_map(self, *args, **kwargs):
        """World coordinates -> map (x, y, z) cell indices. [211-231]"""

    @abc.abstractmethod
    def visualize(self, *args, **kwargs):
        """An image (numpy) summarising the layer for debugging. [233-256]"""
