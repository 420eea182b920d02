"""Scene helpers: where the reference's MJCF files live and the benchmark scenes built on them.

The MJCF files and meshes are the reference's input contract; they are read
from a checkout of hello-robot/stretch_mujoco and are never copied into this repository.
"""
from __future__ import annotations

import os

REFERENCE_MODELS_ENV = "STRETCH_MUJOCO_MODELS"
_DEFAULT_DIRS = ["/root/reference/stretch_mujoco/models"]


def models_dir() -> str | None:
    cand = [os.environ.get(REFERENCE_MODELS_ENV)] + _DEFAULT_DIRS
    try:
        import stretch_mujoco  # the reference package, if installed
        cand.append(os.path.join(os.path.dirname(stretch_mujoco.__file__), "models"))
    except Exception:
        pass
    for c in cand:
        if c and os.path.exists(os.path.join(c, "stretch.xml")):
            return c
    return None


# config 2 of BASELINE.json: stretch.xml + infinite plane + light, i.e. scene.xml minus the dock,
# table and objects.  `lidar=False` drops the <rangefinder> sensors as the
# reference's docs recommend when the lidar is unused (docs/using_mujoco_simulator_with_stretch.md:83).
EMPTY_FLOOR_XML = """<mujoco model="stretch empty floor">
  <include file="stretch.xml"/>
  <statistic center="0 0 .75" extent="1.2" meansize="0.05"/>
  <visual>
    <headlight diffuse="0.6 0.6 0.6" ambient="0.3 0.3 0.3" specular="0 0 0"/>
    <rgba haze="0.15 0.25 0.35 1"/>
  </visual>
  <asset>
    <material name="floor" rgba=".1 .1 .1 1" reflectance="0.1"/>
    <texture type="skybox" builtin="gradient" rgb1="0.44 0.80 1.00" rgb2="1 1 1" width="512" height="3072"/>
  </asset>
  <worldbody>
    <light pos="0 0 1.5" dir="0 0 -1" directional="true"/>
    <geom name="floor" size="0 0 0.05" type="plane" material="floor"/>
  </worldbody>
</mujoco>
"""


def compile_empty_floor(with_render: bool = False):
    from .compiler import compile_scene
    from .mjcf import Scene
    d = models_dir()
    if d is None:
        raise FileNotFoundError("reference MJCF models not found; set $" + REFERENCE_MODELS_ENV)
    return compile_scene(Scene.from_xml_string(EMPTY_FLOOR_XML, base_dir=d), with_render=with_render)


def compile_default_scene(with_render: bool = False):
    from .compiler import compile_scene
    from .mjcf import Scene
    d = models_dir()
    if d is None:
        raise FileNotFoundError("reference MJCF models not found; set $" + REFERENCE_MODELS_ENV)
    return compile_scene(Scene.from_xml_path(os.path.join(d, "scene.xml")), with_render=with_render)


# config 4 of BASELINE.json asks for a Robocasa kitchen, whose assets are download-only and absent here
# (SURVEY.md §2 #13, §8(d)).  KITCHEN PROXY: the box fixtures of the reference's
# third_party/robocasa/.../kitchen_layouts/one_wall_small.yaml -- back / left / right walls (:3-46), floor
# (:48-62), counter_main 2.5 x 0.65 x 0.92 m (:105-113), stove and counter_right to its right -- as MJCF
# boxes around the robot, plus one free box on the counter (nv = 32).  The room frame is shifted so that the
# robot (included at the origin, facing +x) stands 1.1 m in front of the counter.
KITCHEN_PROXY_XML = """<mujoco model="stretch kitchen proxy">
  <include file="stretch.xml"/>
  <statistic center="0 0 .75" extent="1.2" meansize="0.05"/>
  <visual>
    <headlight diffuse="0.6 0.6 0.6" ambient="0.3 0.3 0.3" specular="0 0 0"/>
    <rgba haze="0.15 0.25 0.35 1"/>
  </visual>
  <asset>
    <material name="floor" rgba=".35 .33 .3 1" reflectance="0.1"/>
    <material name="wall" rgba=".85 .83 .78 1"/>
    <material name="counter" rgba=".55 .4 .28 1"/>
    <material name="steel" rgba=".6 .6 .65 1" specular="0.8" shininess="0.6"/>
    <texture type="skybox" builtin="gradient" rgb1="0.44 0.80 1.00" rgb2="1 1 1" width="512" height="3072"/>
  </asset>
  <worldbody>
    <light pos="0 0 2.5" dir="0 0 -1" directional="true"/>
    <geom name="floor" size="0 0 0.05" type="plane" material="floor"/>
    <body name="kitchen" pos="-2.0 1.45 0">
      <geom name="wall" type="box" size="2.75 0.02 1.5" pos="2.75 0.02 1.5" material="wall"/>
      <geom name="wall_left" type="box" size="0.02 1.5 1.5" pos="-0.02 -1.5 1.5" material="wall"/>
      <geom name="wall_right" type="box" size="0.02 1.5 1.5" pos="5.52 -1.5 1.5" material="wall"/>
      <geom name="counter_main" type="box" size="1.25 0.325 0.46" pos="1.5 -0.325 0.46" material="counter"/>
      <geom name="stove" type="box" size="0.38 0.33 0.46" pos="3.13 -0.33 0.46" material="steel"/>
      <geom name="counter_right" type="box" size="0.35 0.325 0.46" pos="3.86 -0.325 0.46" material="counter"/>
      <geom name="sink_tap" type="cylinder" size="0.02 0.12" pos="1.0 -0.12 1.04" material="steel"/>
    </body>
    <body name="obj_box" pos="0.1 1.0 0.97">
      <freejoint/>
      <geom type="box" size=".03 .03 .05" mass=".3" rgba=".8 .2 .2 1"/>
    </body>
  </worldbody>
</mujoco>
"""


def compile_kitchen_proxy(with_render: bool = True, lidar_rays: int = 1000):
    """BASELINE config 4 stand-in; `lidar_rays` re-spins the lidar ring (native 360) at 2*pi/lidar_rays."""
    import math
    from .compiler import compile_scene
    from .mjcf import Scene
    d = models_dir()
    if d is None:
        raise FileNotFoundError("reference MJCF models not found; set $" + REFERENCE_MODELS_ENV)
    ov = None if lidar_rays == 360 else {"lidar": (lidar_rays, "0 0 %.10f" % (2 * math.pi / lidar_rays))}
    return compile_scene(Scene.from_xml_string(KITCHEN_PROXY_XML, base_dir=d, replicate_override=ov), with_render=with_render)
