import os
import sys

import numpy as np
import pytest

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    """gpu-marked tests are skipped (not errored) on a box without CUDA or without the built extension."""
    try:
        import torch
        have = torch.cuda.is_available() and os.path.exists(os.path.join(ROOT, "stretch_mujoco_b200", "libstretchsim.so"))
    except Exception:
        have = False
    if have:
        return
    skip = pytest.mark.skip(reason="needs a CUDA device and the built libstretchsim.so")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


def _blob(name):
    with open(os.path.join(GOLDEN, name), "rb") as fh:
        return fh.read()


@pytest.fixture(scope="session")
def blob_empty_floor():
    return _blob("stretch_empty_floor.ssm")


@pytest.fixture(scope="session")
def blob_default_scene():
    return _blob("stretch_default_scene.ssm")


@pytest.fixture(scope="session")
def oracle_E(blob_empty_floor):
    from oracle.oracle import OracleModel
    return OracleModel(blob_empty_floor)


@pytest.fixture(scope="session")
def arrays_E(blob_empty_floor):
    from stretch_mujoco_b200 import blob
    return blob.unpack(blob_empty_floor)


@pytest.fixture(scope="session")
def settled_home_E(oracle_E, arrays_E):
    """State after 1500 steps (3 s) of the `home` keyframe on the empty floor (oracle, fp64)."""
    A, _ = arrays_E
    om = oracle_E
    qpos = A["qpos0"][None].copy(); qvel = np.zeros((1, om.nv)); warm = np.zeros((1, om.nv)); t = np.zeros(1)
    ctrl = A["key_ctrl"][0][None].copy()
    om.step(qpos, qvel, ctrl, warm, t, nsteps=1500)
    return qpos[0].copy(), qvel[0].copy(), warm[0].copy(), ctrl[0].copy()


def have_reference():
    from stretch_mujoco_b200 import scenes
    return scenes.models_dir() is not None
