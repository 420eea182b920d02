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
