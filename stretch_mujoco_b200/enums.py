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


def get_actuator_by_joint_names_in_mjcf(joint_name: str) -> Actuators:
    """Actuators.get_actuator_by_joint_names_in_mjcf (stretch_mujoco/enums/actuators.py:60-122)."""
    if joint_name == "joint_left_wheel":
        return Actuators.left_wheel_vel
    if joint_name == "joint_right_wheel":
        return Actuators.right_wheel_vel
    if joint_name in ("translate_mobile_base", "position"):
        return Actuators.base_translate
    if joint_name == "rotate_mobile_base":
        return Actuators.base_rotate
    if joint_name == "joint_lift":
        return Actuators.lift
    if "joint_arm" in joint_name:
        return Actuators.arm
    if joint_name == "joint_wrist_yaw":
        return Actuators.wrist_yaw
    if joint_name == "joint_wrist_pitch":
        return Actuators.wrist_pitch
    if joint_name == "joint_wrist_roll":
        return Actuators.wrist_roll
    if joint_name in ("joint_gripper_slide", "gripper_aperture"):
        return Actuators.gripper
    if "joint_gripper_finger_left" in joint_name:
        return Actuators.gripper_left_finger
    if "joint_gripper_finger_right" in joint_name:
        return Actuators.gripper_right_finger
    if joint_name == "joint_head_pan":
        return Actuators.head_pan
    if joint_name == "joint_head_tilt":
        return Actuators.head_tilt
    raise NotImplementedError(f"Actuator for {joint_name} is not defined.")


# slot order of the C-ABI command / status rows (include/stretchsim.h)
COMMAND_SLOTS = ["lift", "arm", "head_pan", "head_tilt", "wrist_yaw", "wrist_pitch", "wrist_roll", "gripper",
                 "base_translate", "base_rotate"]
STATUS_JOINTS = ["lift", "arm", "head_pan", "head_tilt", "wrist_yaw", "wrist_pitch", "wrist_roll", "gripper"]


@dataclass(frozen=True)
class CameraCrop:
    """`CameraCrop` of stretch_mujoco/enums/stretch_cameras.py:161-182."""
    x_min: int
    x_max: int
    y_min: int
    y_max: int

    @property
    def x_offset(self):
        return self.x_min

    @property
    def y_offset(self):
        return self.y_min

    @property
    def width(self):
        return self.x_max - self.x_min

    @property
    def height(self):
        return self.y_max - self.y_min


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
    focal: tuple = (0.0, 0.0)         # calibrated focal lengths (stretch_cameras.py:107-156)
    crop: CameraCrop | None = None
    distortion_params: tuple | None = None

    # the calibration outputs the ROS2 bridge / web teleop consume (stretch_cameras.py:233-295)
    @property
    def field_of_view_vertical_in_degrees(self):
        return self.fovy

    @staticmethod
    def field_of_view_vertical_from_horizontal(fov_horizontal_degrees: int, width: int, height: int) -> int:
        return int(abs(math.degrees(2 * math.atan(math.tan(math.radians(fov_horizontal_degrees) / 2) * width / height))))

    def get_distortion_params_d(self):
        return list(self.distortion_params) if self.distortion_params else [0.0] * 5

    def get_intrinsic_params_k(self):
        return [self.focal[0], 0.0, self.width / 2, 0.0, self.focal[1], self.height / 2, 0.0, 0.0, 1.0]

    def get_projection_matrix_p(self):
        return [self.focal[0], 0.0, self.width / 2, 0.0, 0.0, self.focal[1], self.height / 2, 0.0, 0.0, 0.0, 1.0, 0.0]


def _nav_fovy() -> int:
    # CameraSettings.field_of_view_vertical_from_horizontal(70, 1280, 720): multiplies by the aspect
    # ratio and truncates (stretch_cameras.py:224-231) -> 102
    return int(abs(math.degrees(2 * math.atan(math.tan(math.radians(70) / 2) * 1280 / 720))))


class StretchCameras(Enum):
    cam_d405_rgb = CameraSettings("d405_rgb", 58, 480, 270, False, 0.0, (1280, 720), 0, (242.56, 242.34),
                                  CameraCrop(x_min=125, x_max=395, y_min=0, y_max=270))
    cam_d405_depth = CameraSettings("d405_depth", 58, 480, 270, True, 1.0, (1280, 720), 0, (242.56, 242.34),
                                    CameraCrop(x_min=125, x_max=395, y_min=0, y_max=270))
    cam_d435i_rgb = CameraSettings("d435i_camera_rgb", 42, 424, 240, False, 0.0, (1920, 1080), -1, (304.24, 304.07))
    cam_d435i_depth = CameraSettings("d435i_camera_depth", 42, 424, 240, True, 10.0, (1920, 1080), -1, (304.24, 304.07))
    cam_nav_rgb = CameraSettings("nav_camera_rgb", _nav_fovy(), 800, 600, False, 0.0, (1280, 720), 1, (0.0, 0.0))

    @property
    def camera_name_in_mjcf(self) -> str:
        return self.value.name_in_mjcf

    @property
    def is_depth(self) -> bool:
        return self.value.is_depth

    @property
    def initial_camera_settings(self) -> CameraSettings:
        return self.value

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
