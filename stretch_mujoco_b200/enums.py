"""Names and constants the reference exposes at its API boundary, mirrored for the batched façade.

Sources: `stretch_mujoco/enums/actuators.py:7-26` (actuator names), `stretch_mujoco/enums/
stretch_cameras.py:10-156` (camera names, resolutions, fovy), `stretch_mujoco/enums/
stretch_sensors.py:8-41`, `stretch_mujoco/config.py:1-11`.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from enum import Enum


class Actuators(Enum):
    arm = 0
    gripper = 1
    head_pan = 2
    head_tilt = 3
    lift = 4
    wrist_pitch = 5
    wrist_roll = 6
    wrist_yaw = 7
    base_rotate = 8
    base_translate = 9
    left_wheel_vel = 10
    right_wheel_vel = 11
    gripper_left_finger = 12
    gripper_right_finger = 13


# slot order of the C-ABI command / status rows (include/stretchsim.h)
COMMAND_SLOTS = ["lift", "arm", "head_pan", "head_tilt", "wrist_yaw", "wrist_pitch", "wrist_roll", "gripper",
                 "base_translate", "base_rotate"]
STATUS_JOINTS = ["lift", "arm", "head_pan", "head_tilt", "wrist_yaw", "wrist_pitch", "wrist_roll", "gripper"]


@dataclass(frozen=True)
class CameraSettings:
    name_in_mjcf: str
    fovy: float
    width: int
    height: int
    is_depth: bool
    depth_limit: float = 0.0          # metres; config.depth_limits (config.py:8)
    sensor_resolution: tuple = (0, 0)  # K uses the SENSOR resolution (camera_manager.py:172-182)
    rot90_k: int = 0                  # client-side np.rot90 (status_stretch_camera.py:68-76)


def _nav_fovy() -> int:
    # CameraSettings.field_of_view_vertical_from_horizontal(70, 1280, 720): multiplies by the aspect
    # ratio and truncates (stretch_cameras.py:224-231) -> 102
    return int(abs(math.degrees(2 * math.atan(math.tan(math.radians(70) / 2) * 1280 / 720))))


class StretchCameras(Enum):
    cam_d405_rgb = CameraSettings("d405_rgb", 58, 480, 270, False, 0.0, (1280, 720), 0)
    cam_d405_depth = CameraSettings("d405_depth", 58, 480, 270, True, 1.0, (1280, 720), 0)
    cam_d435i_rgb = CameraSettings("d435i_camera_rgb", 42, 424, 240, False, 0.0, (1920, 1080), -1)
    cam_d435i_depth = CameraSettings("d435i_camera_depth", 42, 424, 240, True, 10.0, (1920, 1080), -1)
    cam_nav_rgb = CameraSettings("nav_camera_rgb", _nav_fovy(), 800, 600, False, 0.0, (1280, 720), 1)

    @staticmethod
    def all():
        return list(StretchCameras)

    @staticmethod
    def none():
        return []

    @staticmethod
    def depth():
        return [c for c in StretchCameras if c.value.is_depth]

    @staticmethod
    def rgb():
        return [c for c in StretchCameras if not c.value.is_depth]


class StretchSensors(Enum):
    base_gyro = 0
    base_accel = 1
    base_lidar = 2

    @staticmethod
    def lidar_names(resolution: int = 360):
        n = len(str(resolution))
        return [f"base_lidar{str(i).zfill(n)}" for i in range(resolution)]


robot_settings = {"wheel_diameter": 0.1016, "wheel_separation": 0.3153, "gripper_min_max": (-0.376, 0.56),
                  "sim_gripper_min_max": (-0.02, 0.04)}
depth_limits = {"d405": 1, "d435i": 10}
base_motion = {"timeout": 15, "default_x_vel": 0.3, "default_r_vel": 1.0}


def compute_K(fovy: float, width: int, height: int):
    """utils.compute_K (stretch_mujoco/utils.py:56-61)."""
    import numpy as np
    f = 0.5 * height / math.tan(fovy * math.pi / 360)
    return np.array(((f, 0, width / 2), (0, f, height / 2), (0, 0, 1)))
