"""Regenerates the compiled-model fixtures from the reference's MJCF files.

Run in a container where /root/reference (hello-robot/stretch_mujoco) is mounted:
    python tests/golden/make_golden.py
The GPU box has no /root/reference, so the -m gpu tests, smoke() and bench.py load these blobs.
"""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
from stretch_mujoco_b200 import blob, scenes  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))

if __name__ == "__main__":
    m = scenes.compile_empty_floor(with_render=False)
    blob.save(os.path.join(HERE, "stretch_empty_floor.ssm"), m)
    print("empty floor:", m.sizes)
    m = scenes.compile_default_scene(with_render=False)
    blob.save(os.path.join(HERE, "stretch_default_scene.ssm"), m)
    print("default scene:", m.sizes)
    # with ray geometry (triangle soups of every ray-visible mesh), zlib-compressed
    m = scenes.compile_empty_floor(with_render=True)
    blob.save(os.path.join(HERE, "stretch_empty_floor_render.ssm.z"), m)
    m = scenes.compile_default_scene(with_render=True)
    blob.save(os.path.join(HERE, "stretch_default_scene_render.ssm.z"), m)
    print("render blobs: %d ray geoms, %d triangles" % (len(m.raygeom_id), len(m.rmesh_face)))
    m = scenes.compile_kitchen_proxy(with_render=True, lidar_rays=1000)
    blob.save(os.path.join(HERE, "stretch_kitchen_proxy_render.ssm.z"), m)
    print("kitchen proxy:", m.sizes)
