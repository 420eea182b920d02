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
from .enums import COMMAND_SLOTS, STATUS_JOINTS, Actuators, StretchCameras, StretchSensors  # noqa: F401


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


class StatusStretchCameras:
    """Field names and accessors of `datamodels/status_stretch_camera.py:10-133`; pixel arrays are torch CUDA
    tensors with a leading env dimension: RGB uint8 [n, H, W, 3], depth float32 [n, H, W], in the camera's own
    orientation and RGB order (what the reference's server produces).  `get_camera_data` applies the client-side
    post-processing of the reference (np.rot90 of the d435i / nav images, RGB->BGR, JET depth colour map) unless
    the frames were already rendered that way (`pull_camera_data(auto_rotate=..., auto_correct_rgb=...)`)."""
    FIELDS = ("cam_d405_rgb", "cam_d405_depth", "cam_d435i_rgb", "cam_d435i_depth", "cam_nav_rgb")

    def __init__(self, time=0.0, fps=0.0, prerotated: bool = False, prebgr: bool = False):
        self.time, self.fps = time, fps
        for n in self.FIELDS:
            setattr(self, n, None)
        self.cam_d405_K = None
        self.cam_d435i_K = None
        self._prerotated, self._prebgr = prerotated, prebgr

    @staticmethod
    def default():
        return StatusStretchCameras(time=0, fps=0)

    def set_camera_data(self, camera, data) -> None:
        if camera.name not in self.FIELDS:
            raise NotImplementedError(f"Camera {camera} is not implemented.")
        setattr(self, camera.name, data)

    def get_camera_data(self, camera, *, auto_rotate: bool = True, auto_correct_rgb: bool = True,
                        use_depth_color_map: bool = False):
        import torch
        data = getattr(self, camera.name, None) if camera.name in self.FIELDS else None
        if data is None:
            raise ValueError(f"Tried to get {camera} data, but it is empty or not implemented.")
        k = camera.value.rot90_k
        if auto_rotate and k and not self._prerotated:
            data = torch.rot90(data, k, dims=(1, 2))          # np.rot90(img, k) per env
        if camera.value.is_depth:
            if use_depth_color_map:
                data = engine.depth_colormap(data)            # utils.get_depth_color_map (utils.py:363-373)
        elif auto_correct_rgb and not self._prebgr:
            data = data.flip(-1)                              # cv2.COLOR_RGB2BGR
        return data

    def get_all(self, *, auto_rotate: bool = True, auto_correct_rgb: bool = True, use_depth_color_map: bool = False) -> dict:
        out = {}
        for cam in StretchCameras.all():
            try:
                out[cam] = self.get_camera_data(cam, auto_rotate=auto_rotate, auto_correct_rgb=auto_correct_rgb,
                                                use_depth_color_map=use_depth_color_map)
            except ValueError:
                pass
        return out

    def to_dict(self) -> dict:
        d = {"time": self.time, "fps": self.fps, "cam_d405_K": self.cam_d405_K, "cam_d435i_K": self.cam_d435i_K}
        d.update({n: getattr(self, n) for n in self.FIELDS})
        return d

    # dict-style access by StretchCameras member or field name (raw stored frames)
    def __getitem__(self, key):
        name = key if isinstance(key, str) else key.name
        v = getattr(self, name, None)
        if v is None:
            raise KeyError(key)
        return v

    def __contains__(self, key):
        name = key if isinstance(key, str) else key.name
        return getattr(self, name, None) is not None


class StatusStretchSensors:
    """Field names and accessors of `datamodels/status_stretch_sensors.py:11-77`; values are [nenv, ...] tensors."""

    def __init__(self, time=0.0, fps=0.0, base_gyro=None, base_imu=None, lidar=None):
        self.time, self.fps, self.base_gyro, self.base_imu, self.lidar = time, fps, base_gyro, base_imu, lidar

    @staticmethod
    def default():
        return StatusStretchSensors(time=0, fps=0)

    def get_data(self, sensor):
        data = {StretchSensors.base_gyro: self.base_gyro, StretchSensors.base_accel: self.base_imu,
                StretchSensors.base_lidar: self.lidar}.get(sensor)
        if data is None:
            raise ValueError(f"Tried to get {sensor} data, but it is empty.")
        return data

    def set_data(self, sensor, value) -> None:
        if sensor == StretchSensors.base_gyro:
            self.base_gyro = value
        elif sensor == StretchSensors.base_accel:
            self.base_imu = value
        elif sensor == StretchSensors.base_lidar:
            self.lidar = value
        else:
            raise NotImplementedError(f"Sensor {sensor} is not implemented")

    def to_dict(self) -> dict:
        return {"time": self.time, "fps": self.fps, "base_gyro": self.base_gyro, "base_imu": self.base_imu, "lidar": self.lidar}

    def __getitem__(self, sensor):
        return self.get_data(sensor)


class StretchMujocoSimulator:
    def __init__(self, scene_xml_path: str | None = None, model=None, camera_hz: float = 30,
                 cameras_to_use: list | None = None, start_translation=None, start_rotation_quat=None, *,
                 nenv: int = 1, device: int = 0, model_blob: bytes | None = None, maxcon: int = 32, maxefc: int = 0,
                 with_render: bool | None = None):
        self.scene_xml_path, self.model_blob = scene_xml_path, model_blob
        self.camera_hz = camera_hz
        self.cameras_to_use = list(cameras_to_use or [])
        self.start_translation, self.start_rotation_quat = start_translation, start_rotation_quat
        self.nenv, self.device, self.maxcon, self.maxefc = nenv, device, maxcon, maxefc
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
        self.batch = engine.Batch(self.dmodel, self.nenv, maxcon=self.maxcon, maxefc=self.maxefc)
        self._act = {n: self.dmodel.name2id(engine.OBJ_ACTUATOR, n) for n in
                     ["lift", "arm", "head_pan", "head_tilt", "wrist_yaw", "wrist_pitch", "wrist_roll", "gripper",
                      "left_wheel_vel", "right_wheel_vel"]}
        self._last_move_to: dict[str, object] = {}   # per actuator: [nenv] tensor of the last move_to target (NaN = none)
        self._base_busy, self._base_checked = False, 0
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
        """Apply pending commands (P2), advance `nsteps` physics steps.

        The reference runs push_command + BaseController.update after EVERY mj_step
        (mujoco_server.py:378-379,450-463).  Commands are edge-triggered, so for everything except a running
        base move_by one application in front of the block of steps is equivalent; while a base
        translate_by / rotate_by may be running (its stop test must be evaluated per step) the block is
        stepped in the reference's cadence, one command pass per physics step."""
        self._require()
        if self._base_busy:
            self.batch.step_controlled(nsteps)
            self._base_checked += nsteps
            if self._base_checked >= 50:          # one small D2H read per 50 steps to leave the slow cadence
                self._base_checked = 0
                mode = self.batch.base_state[:, 0]
                self._base_busy = bool(((mode == 1) | (mode == 2)).any())
            return
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
        import torch
        tgt = self._last_move_to.get(a.name)
        if tgt is None:
            tgt = torch.full((self.nenv,), float("nan"), dtype=torch.float32, device=self.batch.qpos.device)
            self._last_move_to[a.name] = tgt
        tgt[slice(None) if env_ids is None else env_ids] = torch.as_tensor(pos, dtype=torch.float32, device=tgt.device)

    def move_by(self, actuator, pos, env_ids=None) -> None:
        self._require()
        a = self._slot(actuator)
        if a in (Actuators.left_wheel_vel, Actuators.right_wheel_vel):
            raise Exception(f"Cannot set an absolute position for a continuous joint {a.name}")
        if a.name not in COMMAND_SLOTS:
            raise NotImplementedError(f"Actuator {a.name} is not supported.")
        self._write(20, COMMAND_SLOTS.index(a.name), pos, env_ids)
        if a in (Actuators.base_translate, Actuators.base_rotate):
            self._base_busy, self._base_checked = True, 0

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
        """{Actuators: (min, max)} like MujocoServer.update_joint_limits (mujoco_server.py:281-291): every MJCF
        joint that maps to an actuator (enums/actuators.py:60-122); later joints of the same actuator (the four
        telescope segments) overwrite earlier ones, as in the reference."""
        self._require()
        rng = self.dmodel.get("jnt_range").reshape(-1, 2)
        out = {}
        for j in range(self.dmodel.njnt):
            try:
                act = enums.get_actuator_by_joint_names_in_mjcf(self.dmodel.id2name(engine.OBJ_JOINT, j))
            except NotImplementedError:
                continue
            out[act] = (float(rng[j, 0]), float(rng[j, 1]))
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

    def pull_sensor_data(self) -> StatusStretchSensors:
        """gyro [nenv,3], accelerometer [nenv,3], lidar [nenv,nray] as a `StatusStretchSensors`
        (status_stretch_sensors.py:11-77; gathered like mujoco_server_sensor_manager.py:65-89)."""
        self._require()
        out = StatusStretchSensors(time=self.batch.time, fps=0.0, base_gyro=self.batch.sensordata[:, 0:3],
                                   base_imu=self.batch.sensordata[:, 3:6])
        if self.dmodel.nrange > 0:
            out.lidar = self.batch.lidar()
        return out

    def pull_camera_data(self, cameras: list | None = None, width: int | None = None, height: int | None = None,
                         env_begin: int = 0, env_count: int | None = None, auto_rotate: bool = False,
                         auto_correct_rgb: bool = False) -> StatusStretchCameras:
        """`StatusStretchCameras` (indexable by StretchCameras member): RGB uint8 [n,H,W,3] or depth float32 [n,H,W] with the depth
        limit applied (camera_manager.py:127-143, utils.py:87-91).  By default images are in camera
        orientation, RGB order (what the reference's server produces).  `auto_rotate` / `auto_correct_rgb`
        apply what StatusStretchCameras.get_camera_data does on the client
        (datamodels/status_stretch_camera.py:47-82): np.rot90(img, -1) for the d435i cameras, np.rot90(img, 1)
        for the nav camera (outputs become [n,W,H(,3)]) and RGB->BGR -- fused into the render kernel."""
        self._require()
        import torch
        cams = cameras if cameras is not None else self.cameras_to_use
        n = self.nenv - env_begin if env_count is None else env_count
        out = StatusStretchCameras(time=self.batch.time, fps=0.0, prerotated=auto_rotate, prebgr=auto_correct_rgb)
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
            out.set_camera_data(cam, img)
        out.cam_d405_K = enums.compute_K(58, 1280, 720)    # camera_manager.py:168-183
        out.cam_d435i_K = enums.compute_K(42, 1920, 1080)
        return out

    # ------------------------------------------------------------------ blocking helpers
    def wait_until_at_setpoint(self, actuator, timeout: float = 5.0, position_tolerance: float = 0.05) -> bool:
        """Steps the batch until every env is within tolerance of the last move_to, or `timeout`
        seconds of SIMULATED time pass (stretch_mujoco_simulator.py:267-297)."""
        import torch
        a = self._slot(actuator)
        if a.name not in self._last_move_to:
            return True
        target = self._last_move_to[a.name]           # [nenv]; NaN where this env never got a move_to for the actuator
        dt = float(self.dmodel.get("opt_timestep").ravel()[0])
        for _ in range(int(timeout / (dt * 50)) + 1):
            self.step(50)
            pos = getattr(self.pull_status(), a.name).pos
            if bool(torch.all(torch.isnan(target) | (torch.abs(pos - target) <= position_tolerance))):
                return True
        return False

    def wait_while_is_moving(self, actuator, timeout: float = 5.0, check_interval: float = 0.1,
                             position_tolerance: float = 1e-4) -> bool:
        import torch
        a = self._slot(actuator)
        dt = float(self.dmodel.get("opt_timestep").ravel()[0])
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
