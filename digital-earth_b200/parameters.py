"""Host mirrors of lib/parameters.py:4-15 (field names and meaning kept)."""
from dataclasses import dataclass, field
from typing import Tuple

Vec3 = Tuple[float, float, float]


@dataclass
class PathParameters:  # lib/parameters.py:4-8
    wavelength: float = 0.0
    ray_dir: Vec3 = (0.0, 0.0, 0.0)
    ray_pos: Vec3 = (0.0, 0.0, 0.0)


@dataclass
class SceneParameters:  # lib/parameters.py:10-15
    light_direction: Vec3 = field(default=(0.0, 0.0, 0.0))
    sun_cos_angle: float = 0.0
    sun_angular_radius: float = 0.0
    land_height_scale: float = 0.0
