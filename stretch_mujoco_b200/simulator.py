"""Batched drop-in façade with the reference's `StretchMujocoSimulator` method names
(`stretch_mujoco/stretch_mujoco_simulator.py:34-534`): start / stop / home / stow / move_to /
move_by / set_base_velocity / pull_status / pull_camera_data / pull_sensor_data / ...

Differences that follow from batching: every value has a leading `nenv`
dimension and is a torch CUDA tensor; commands take an optional `env_ids`; there is no server
process or real-time sleep, so simulated time advances only inside `step()` (and inside the
`wait_*` helpers, which step until their condition holds).
"""
from __future__ import annotations

import os

import numpy as np

from . import engine, enums
from .enums import COMMAND_SLOTS, STATUS_JOINTS, Actuators, StretchCameras, StretchSensors


class PositionVelocity:
    def __init__(self, pos, vel):
        self.pos, self.vel = pos, vel


class BaseStatus:
    def __init__(self, x, y, theta, x_vel, theta_vel):
        self.x, self.y, self.theta, self.x_vel, self.theta_vel = x, y, theta, x_vel, theta_vel


class StatusStretchJoints:
    """Field names of `datamodels/status_stretch_joints.py:27-74`; values are [nenv] tensors."""

    def __init__(self, row):
        self.raw = row
        self.time = row[:, 0]
        for k, n in enumerate(STATUS_JOINTS):
            setattr(self, n, PositionVelocity(row[:, 1 + 2 * k], row[:, 2 + 2 * k]))
        self.base = BaseStatus(row[:, 17], row[:, 18], row[:, 19], row[:, 20], row[:, 21])


class StretchMujocoSimulator:
    def __init__(self, scene_xml_path: str | None = None, model=None, camera_hz: float = 30,
                 cameras_to_use: list | None = None, start_translation=None, start_rotation_quat=None, *,
                 nenv: int = 1, device: int = 0, model_blob: bytes | None = None, maxcon: int = 24,
                 with_render: bool | None = None):
        self.scene_xml_path, self.model_blob = scene_xml_path, model_blob
        self.camera_hz = camera_hz
        self.cameras_to_use = list(cameras_to_use or [])
        self.start_translation, self.start_rotation_quat = start_translation, start_rotation_quat
        self.nenv, self.device, self.maxcon = nenv, device, maxcon
        self.with_render = bool(self.cameras_to_use) if with_render is None else with_render
        self._running = False
        self.batch = None
        self.dmodel = model

    # ------------------------------------------------------------------ lifecycle
    def start(self, show_viewer_ui: bool = False, headless: bool = True, use_passive_viewer: bool = True,
              home: bool = True) -> None:
        if show_viewer_ui or not headless:
            raise NotImplementedError("the batched engine is headless (viewer servers are out of scope)")
        if self.dmodel is None:
            blob_bytes = self.model_blob
            if blob_bytes is None:
                from . import blob, scenes
                from .compiler import compile_scene
                from .mjcf import Scene
                path = self.scene_xml_path or os.path.join(scenes.models_dir() or "", "scene.xml")
                m = compile_scene(Scene.from_xml_path(path), with_render=self.with_render)
                blob_bytes = blob.pack(m.arrays, m.names)
            self.dmodel = engine.DeviceModel(blob_bytes, self.device)
        if self.start_translation is not None or self.start_rotation_quat is not None:
            # change_start_pose (stretch_mujoco/mujoco_server.py:206-229): edit qpos0 of base_link's free joint
            q0 = self.dmodel.get("qpos0")
            if self.start_translation is not None:
                q0[0:3] = self.start_translation
            if self.start_rotation_quat is not None:
                q0[3:7] = self.start_rotation_quat
            self.dmodel.set("qpos0", q0)
        self.batch = engine.Batch(self.dmodel, self.nenv, maxcon=self.maxcon)
        self._act = {n: self.dmodel.name2id(engine.OBJ_ACTUATOR, n) for n in
                     ["lift", "arm", "head_pan", "head_tilt", "wrist_yaw", "wrist_pitch", "wrist_roll", "gripper",
                      "left_wheel_vel", "right_wheel_vel"]}
        self._last_move_to: dict[str, object] = {}
        self._running = True
        self.batch.forward()
        if home:
            self.home()

    def stop(self) -> None:
        self._running = False
        self.batch = None

    def is_running(self) -> bool:
        return self._running

    def _require(self):
        if not self._running:
            raise ConnectionError("The Stretch Mujoco Simulator is not running. Call start() first.")  # utils.py:43-53

    # ------------------------------------------------------------------ stepping
    def step(self, nsteps: int = 1) -> None:
        """Apply pending commands (P2), advance `nsteps` physics steps."""
        self._require()
        self.batch.apply_commands()
        self.batch.step(nsteps)

    # ------------------------------------------------------------------ commands
    def _slot(self, actuator):
        if isinstance(actuator, str):
            actuator = Actuators[actuator]
        return actuator

    def _write(self, base: int, slot: int, pos, env_ids):
        import torch
        c = self.batch.command
        pos_t = torch.as_tensor(pos, dtype=torch.float32, device=c.device)
        if env_ids is None:
            c[:, base + slot] = 1.0
            c[:, base + 10 + slot] = pos_t
        else:
            c[env_ids, base + slot] = 1.0
            c[env_ids, base + 10 + slot] = pos_t

    def move_to(self, actuator, pos, env_ids=None) -> None:
        self._require()
        a = self._slot(actuator)
        if a in (Actuators.left_wheel_vel, Actuators.right_wheel_vel, Actuators.base_rotate, Actuators.base_translate):
            raise Exception(f"Cannot set an absolute position for a continuous joint {a.name}")
        if a.name not in COMMAND_SLOTS:
            raise NotImplementedError(f"Actuator {a.name} is not supported.")
        self._write(0, COMMAND_SLOTS.index(a.name), pos, env_ids)
        self._last_move_to[a.name] = pos

    def move_by(self, actuator, pos, env_ids=None) -> None:
        self._require()
        a = self._slot(actuator)
        if a in (Actuators.left_wheel_vel, Actuators.right_wheel_vel):
            raise Exception(f"Cannot set an absolute position for a continuous joint {a.name}")
        if a.name not in COMMAND_SLOTS:
            raise NotImplementedError(f"Actuator {a.name} is not supported.")
        self._write(20, COMMAND_SLOTS.index(a.name), pos, env_ids)

    def set_base_velocity(self, v_linear, omega, env_ids=None) -> None:
        self._require()
        import torch
        c = self.batch.command
        sel = slice(None) if env_ids is None else env_ids
        c[sel, 40] = 1.0
        c[sel, 41] = torch.as_tensor(v_linear, dtype=torch.float32, device=c.device)
        c[sel, 42] = torch.as_tensor(omega, dtype=torch.float32, device=c.device)

    def _keyframe(self, name: str, env_ids=None):
        self._require()
        k = self.dmodel.name2id(engine.OBJ_KEY, name)
        if k < 0:
            raise ValueError(f"model has no keyframe '{name}'")
        sel = slice(None) if env_ids is None else env_ids
        self.batch.command[sel, 43] = float(k + 1)

    def home(self, env_ids=None) -> None:
        self._keyframe("home", env_ids)

    def stow(self, env_ids=None) -> None:
        self._keyframe("stow", env_ids)

    def set_ctrl(self, ctrl) -> None:
        """Raw actuator targets [nenv, nu] (device tensor or pinned host tensor): the data-parallel
        entry used by rollouts; bypasses the edge-triggered command slots."""
        self._require()
        self.batch.ctrl.copy_(ctrl, non_blocking=True)

    # ------------------------------------------------------------------ observations
    def pull_status(self) -> StatusStretchJoints:
        self._require()
        return StatusStretchJoints(self.batch.pull_status())

    def pull_joint_limits(self) -> dict:
        """{actuator name: (min, max)} like mujoco_server.update_joint_limits (mujoco_server.py:281-291)."""
        self._require()
        rng = self.dmodel.get("jnt_range").reshape(-1, 2)
        out = {}
        for j in range(self.dmodel.njnt):
            out[self.dmodel.id2name(engine.OBJ_JOINT, j)] = (float(rng[j, 0]), float(rng[j, 1]))
        return out

    def get_base_pose(self):
        s = self.pull_status()
        return s.base.x, s.base.y, s.base.theta

    def get_link_pose(self, link_name: str):
        """World position [nenv,3] and quaternion [nenv,4] of a body (the reference goes through a
        URDF FK, stretch_mujoco_simulator.py:468-486; here the engine's own body frames are used)."""
        self._require()
        b = self.dmodel.name2id(engine.OBJ_BODY, link_name)
        if b < 0:
            raise ValueError(f"unknown link {link_name}")
        return self.batch.xpos[:, b], self.batch.xquat[:, b]

    def get_ee_pose(self):
        return self.get_link_pose("link_grasp_center")

    def pull_sensor_data(self) -> dict:
        """gyro [nenv,3], accelerometer [nenv,3], lidar [nenv,nray] (status_stretch_sensors.py:11-77)."""
        self._require()
        out = {StretchSensors.base_gyro: self.batch.sensordata[:, 0:3], StretchSensors.base_accel: self.batch.sensordata[:, 3:6]}
        if self.dmodel.nrange > 0:
            out[StretchSensors.base_lidar] = self.batch.lidar()
        return out

    def pull_camera_data(self, cameras: list | None = None, width: int | None = None, height: int | None = None,
                         env_begin: int = 0, env_count: int | None = None, auto_rotate: bool = False,
                         auto_correct_rgb: bool = False) -> dict:
        """{StretchCameras: tensor}: RGB uint8 [n,H,W,3] or depth float32 [n,H,W] with the depth
        limit applied (camera_manager.py:127-143, utils.py:87-91).  By default images are in camera
        orientation, RGB order (what the reference's server produces).  `auto_rotate` / `auto_correct_rgb`
        apply what StatusStretchCameras.get_camera_data does on the client
        (datamodels/status_stretch_camera.py:47-82): np.rot90(img, -1) for the d435i cameras, np.rot90(img, 1)
        for the nav camera (outputs become [n,W,H(,3)]) and RGB->BGR -- fused into the render kernel."""
        self._require()
        import torch
        cams = cameras if cameras is not None else self.cameras_to_use
        n = self.nenv - env_begin if env_count is None else env_count
        out = {}
        for cam in cams:
            cs = cam.value
            cid = self.dmodel.name2id(engine.OBJ_CAMERA, cs.name_in_mjcf)
            if cid < 0:
                raise ValueError(f"Tried to get {cam} imagery, but it is not available")
            W, H = width or cs.width, height or cs.height
            dev = self.batch.qpos.device
            rot = 0
            if auto_rotate:
                rot = -1 if "d435i" in cs.name_in_mjcf else (1 if "nav" in cs.name_in_mjcf else 0)
            oh, ow = (W, H) if rot else (H, W)
            if cs.is_depth:
                img = torch.empty(n, oh, ow, dtype=torch.float32, device=dev)
                self.batch.render(cid, W, H, cs.fovy, None, img, cs.depth_limit, env_begin, n, rot90=rot)
            else:
                img = torch.empty(n, oh, ow, 3, dtype=torch.uint8, device=dev)
                self.batch.render(cid, W, H, cs.fovy, img, None, 0.0, env_begin, n, rot90=rot, bgr=auto_correct_rgb)
            out[cam] = img
        out["cam_d405_K"] = enums.compute_K(58, 1280, 720)    # camera_manager.py:168-183
        out["cam_d435i_K"] = enums.compute_K(42, 1920, 1080)
        return out

    # ------------------------------------------------------------------ blocking helpers
    def wait_until_at_setpoint(self, actuator, timeout: float = 5.0, position_tolerance: float = 0.05) -> bool:
        """Steps the batch until every env is within tolerance of the last move_to, or `timeout`
        seconds of SIMULATED time pass (stretch_mujoco_simulator.py:267-297)."""
        import torch
        a = self._slot(actuator)
        if a.name not in self._last_move_to:
            return True
        target = torch.as_tensor(self._last_move_to[a.name], dtype=torch.float32, device=self.batch.qpos.device)
        dt = float(self.dmodel.get("opt_timestep")[0])
        for _ in range(int(timeout / (dt * 50)) + 1):
            self.step(50)
            pos = getattr(self.pull_status(), a.name).pos
            if bool(torch.all(torch.abs(pos - target) <= position_tolerance)):
                return True
        return False

    def wait_while_is_moving(self, actuator, timeout: float = 5.0, check_interval: float = 0.1,
                             position_tolerance: float = 1e-4) -> bool:
        import torch
        a = self._slot(actuator)
        dt = float(self.dmodel.get("opt_timestep")[0])
        n = max(int(check_interval / dt), 1)
        last = None
        for _ in range(int(timeout / check_interval) + 1):
            self.step(n)
            s = self.pull_status()
            cur = torch.stack([s.base.x, s.base.y, s.base.theta], 1) if a in (Actuators.base_rotate, Actuators.base_translate) \
                else getattr(s, a.name).pos.clone()
            if last is not None and bool(torch.all(torch.abs(cur - last) <= position_tolerance)):
                return True
            last = cur.clone()
        return False
