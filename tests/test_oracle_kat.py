"""Pins the CPU oracle.  CPU only.

The reference's tests hold no golden vectors; the anchors that exist are numbers the reference
itself PRINTED after running its own MuJoCo path: `docs/getting_started.ipynb:729-751` (status at
t = 8.26 s after `home`), `:799` (head-tilt limit saturation), `README.md:136-145`.  The oracle
reproduces them to 4-8 digits, which pins the loader's mass properties, gravcomp, the actuator
model, friction loss, joint limits, equality constraints and the Newton solve end to end.
"""
import numpy as np
import pytest

from stretch_mujoco_b200 import compiler


def _rollout(om, A, nsteps, ctrl=None, want=("act_length", "act_velocity")):
    qpos = A["qpos0"][None].copy(); qvel = np.zeros((1, om.nv)); warm = np.zeros((1, om.nv)); t = np.zeros(1)
    ctrl = A["key_ctrl"][0][None].copy() if ctrl is None else ctrl
    out = om.step(qpos, qvel, ctrl, warm, t, nsteps=nsteps, want=want)
    return qpos, qvel, warm, t, out


def test_home_status_matches_reference_notebook(oracle_E, arrays_E):
    A, _ = arrays_E
    qpos, qvel, warm, t, o = _rollout(oracle_E, A, 4130)  # t = 8.26 s
    L, V = o["act_length"][0], o["act_velocity"][0]
    ref = dict(lift=(0.5905520090306994, 2.2063552289719744e-4), arm=(0.09999622635034094, 1.34e-10),
               head_pan=(-5.005046374741913e-06, 0), head_tilt=(-0.004519272499335126, 0),
               wrist_yaw=(9.232975816659571e-05, -3.3150607193177336e-05), wrist_pitch=(-0.005324523093874352, 0),
               wrist_roll=(-9.586627571896982e-05, 0))
    idx = dict(lift=2, arm=3, wrist_yaw=4, wrist_pitch=5, wrist_roll=6, head_pan=8, head_tilt=9)
    assert t[0] == pytest.approx(8.26, abs=1e-9)
    assert L[idx["lift"]] == pytest.approx(ref["lift"][0], abs=2e-5)       # mid-creep value, 5 digits
    assert V[idx["lift"]] == pytest.approx(ref["lift"][1], rel=2e-2)
    assert L[idx["arm"]] == pytest.approx(ref["arm"][0], abs=1e-8)
    assert L[idx["head_tilt"]] == pytest.approx(ref["head_tilt"][0], abs=1e-8)
    assert L[idx["head_pan"]] == pytest.approx(ref["head_pan"][0], abs=1e-8)
    assert L[idx["wrist_pitch"]] == pytest.approx(ref["wrist_pitch"][0], abs=1e-8)
    assert L[idx["wrist_roll"]] == pytest.approx(ref["wrist_roll"][0], abs=5e-7)
    assert L[idx["wrist_yaw"]] == pytest.approx(ref["wrist_yaw"][0], abs=2e-5)
    # gripper: sim 0 maps to -0.0639975 in the real range (config.py:4-5, mujoco_server.py:517-525)
    g = (L[7] + 0.02) * (0.56 + 0.376) / 0.06 - 0.376
    assert g == pytest.approx(-0.06399746756801022, abs=1e-5)


def test_head_tilt_limit_saturation(oracle_E, arrays_E):
    A, _ = arrays_E
    om = oracle_E
    qpos, qvel, warm, t, _ = _rollout(om, A, 2000)
    ctrl = A["key_ctrl"][0][None].copy(); ctrl[0, 9] = -2.0  # move_to('head_tilt', -2.0)
    o = om.step(qpos, qvel, ctrl, warm, t, nsteps=3000, want=("act_length",))
    assert o["act_length"][0, 9] == pytest.approx(-1.52257, abs=2e-5)  # docs/getting_started.ipynb:799


def test_mass_matrix_matches_independent_jacobian_form(oracle_E, blob_empty_floor):
    from stretch_mujoco_b200 import blob
    m = blob.load.__globals__["unpack"](blob_empty_floor)
    model = compiler.Model(); model.arrays, model.names = m
    rng = np.random.default_rng(1)
    q = model.qpos0.copy()
    q[7:] += rng.uniform(-0.1, 0.1, size=len(q) - 7)
    quat = rng.normal(size=4); q[3:7] = quat / np.linalg.norm(quat)
    M2 = compiler.mass_matrix_numpy(model, q)[0]
    o = oracle_E.forward(q[None], np.zeros((1, oracle_E.nv)), np.zeros((1, oracle_E.nu)), want=("M",))
    assert np.allclose(o["M"][0], M2, rtol=0, atol=1e-10 * np.abs(M2).max())
    assert np.all(np.linalg.eigvalsh(M2) > 0)


def test_wheel_and_caster_contacts_at_rest(oracle_E, arrays_E, settled_home_E):
    A, names = arrays_E
    qpos, qvel, warm, ctrl = settled_home_E
    o = oracle_E.forward(qpos[None], qvel[None], ctrl[None], warm[None], want=("ncon", "contact_geom", "nefc", "contact_dist"))
    gname = names[compiler.OBJ_GEOM]
    assert o["ncon"][0] == 5  # caster sphere + 2 per wheel cylinder (plane-cylinder rim points)
    types = sorted(int(A["geom_type"][g]) for g in o["contact_geom"][0, :5, 1])
    assert types == [2, 5, 5, 5, 5] and all(gname[g] == "floor" for g in o["contact_geom"][0, :5, 0])
    assert o["nefc"][0] == 5 + 12 + 0 + 1 + 4 * 6  # equality + friction loss + limits + caster + condim-6 wheels
    assert np.all(o["contact_dist"][0, :5] < 0) and np.all(o["contact_dist"][0, :5] > -2e-3)


def test_free_fall_and_accelerometer(blob_default_scene):
    """object1 of scene.xml is a free box dropped from z=0.6: analytic free fall until it lands."""
    from oracle.oracle import OracleModel
    from stretch_mujoco_b200 import blob
    A, names = blob.unpack(blob_default_scene)
    om = OracleModel(blob_default_scene)
    qpos = A["qpos0"][None].copy(); qvel = np.zeros((1, om.nv)); warm = np.zeros((1, om.nv)); t = np.zeros(1)
    ctrl = A["key_ctrl"][0][None].copy()
    z0 = qpos[0, 34 + 2]
    n = 50
    om.step(qpos, qvel, ctrl, warm, t, nsteps=n)
    h = 0.002
    # semi-implicit Euler: v_k = -g h k ; z_n = z0 - g h^2 n(n+1)/2
    assert qvel[0, 32 + 2] == pytest.approx(-9.81 * h * n, rel=1e-9)
    assert qpos[0, 34 + 2] == pytest.approx(z0 - 9.81 * h * h * n * (n + 1) / 2, abs=1e-12)


def test_imu_at_rest(oracle_E, settled_home_E):
    qpos, qvel, warm, ctrl = settled_home_E
    o = oracle_E.forward(qpos[None], qvel[None], ctrl[None], warm[None], want=("sensordata",))
    gyro, acc = o["sensordata"][0, :3], o["sensordata"][0, 3:6]
    assert np.abs(gyro).max() < 1e-2
    # IMU is mounted upside-down: reads ~(0, 0, -9.81) at rest
    assert acc[2] == pytest.approx(-9.81, abs=0.1) and np.abs(acc[:2]).max() < 0.2


def test_solver_optimality(oracle_E, settled_home_E):
    """At the solution M*qacc = qfrc_smooth + J^T f (stationarity of the Newton objective)."""
    qpos, qvel, warm, ctrl = settled_home_E
    rng = np.random.default_rng(3)
    qv = qvel + rng.normal(scale=0.05, size=qvel.shape)
    o = oracle_E.forward(qpos[None], qv[None], ctrl[None], None, maxefc=96,
                         want=("M", "qacc", "qfrc_constraint", "qacc_smooth", "efc_J", "efc_force", "nefc"))
    M, qacc = o["M"][0], o["qacc"][0]
    lhs = M @ (qacc - o["qacc_smooth"][0])
    assert np.allclose(lhs, o["qfrc_constraint"][0], atol=1e-6 * max(1.0, np.abs(lhs).max()))
    n = o["nefc"][0]
    assert np.allclose(o["efc_J"][0, :n].T @ o["efc_force"][0, :n], o["qfrc_constraint"][0], atol=1e-9)


def test_solver_optimality_on_random_rollout_states(oracle_E, arrays_E):
    """Regression for the line-search defect the device comparison exposed: on states of the
    bench workload (limits, self-contact, sliding wheels) the Newton solve must end at a stationary point,
    M qacc - qfrc_smooth - J^T f = 0, in EVERY env -- a premature stop showed up as residuals of 200-650."""
    import bench
    A, _ = arrays_E
    om = oracle_E
    om.set_options(enable_lidar=False)
    nenv = 256
    lo, hi = A["actuator_ctrlrange"][:, 0].copy(), A["actuator_ctrlrange"][:, 1].copy()
    q = np.tile(A["qpos0"], (nenv, 1)); v = np.zeros((nenv, om.nv)); w = np.zeros((nenv, om.nv)); t = np.zeros(nenv)
    worst = 0.0
    for p in range(3):
        c = bench.ctrl_np(0, 0, nenv, p, lo, hi)
        for chunk in (7, 43):
            om.step(q, v, c, w, t, nsteps=chunk)
            o = om.forward(q, v, c, w, want=("qacc", "M", "qfrc_bias", "qfrc_passive", "qfrc_actuator", "qfrc_constraint", "solver_iter", "flags"))
            smooth = o["qfrc_passive"] - o["qfrc_bias"] + o["qfrc_actuator"]
            r = np.einsum("eij,ej->ei", o["M"], o["qacc"]) - smooth - o["qfrc_constraint"]
            scale = np.maximum(np.abs(smooth).max(1), 1.0)
            worst = max(worst, float((np.abs(r).max(1) / scale).max()))
            assert o["solver_iter"].max() < 50
    assert worst < 1e-5, worst


def test_ray_primitives_closed_form(blob_default_scene):
    """Pins the oracle's mj_ray restatement (row S2) on closed-form ray / primitive intersections in the
    reference's default scene (models/scene.xml:21-35): floor plane, the table box (pos 0 -1 .24, half sizes
    .6 .5 .24), the free cylinder object2 (radius .02, half height .04 at 0.08 -0.55 0.6) and a miss."""
    import os
    from oracle.oracle import OracleModel
    from stretch_mujoco_b200 import blob
    raw = blob.read_bytes(os.path.join(os.path.dirname(__file__), "golden", "stretch_default_scene_render.ssm.z"))   # with ray geometry
    A, names = blob.unpack(raw)
    om = OracleModel(raw)
    om.set_options(enable_lidar=False)
    q = A["qpos0"][None].copy()
    o = om.forward(q, np.zeros((1, om.nv)), np.zeros((1, om.nu)), want=("xpos", "xquat"))
    gname = names[compiler.OBJ_GEOM]
    body = names[compiler.OBJ_BODY]
    gb = A["geom_bodyid"]
    org = np.array([[[5.0, 5.0, 2.0], [0.3, -1.2, 2.0], [2.0, -1.2, 0.24], [0.08, -0.55, 2.0], [5.0, 5.0, 2.0], [0.3, -1.2, 2.0]]])
    drc = np.array([[[0, 0, -1.0], [0, 0, -1.0], [-1.0, 0, 0], [0, 0, -1.0], [0, 0, 1.0], [0, 0.28, -0.96]]])
    dist, geom = om.rays(o["xpos"], o["xquat"], org, drc, groupmask=0, bodyexclude=-1)
    d, g = dist[0], geom[0]
    assert d[0] == pytest.approx(2.0, abs=1e-12) and gname[g[0]] == "floor"
    assert d[1] == pytest.approx(2.0 - 0.48, abs=1e-12) and body[gb[g[1]]] == "table"          # top face of the table
    assert d[2] == pytest.approx(2.0 - 0.6, abs=1e-12) and body[gb[g[2]]] == "table"           # +x side face
    assert d[3] == pytest.approx(2.0 - 0.64, abs=1e-12) and body[gb[g[3]]] == "object2"        # top cap of the cylinder
    assert d[4] == -1.0 and g[4] == -1                                                         # looking at the sky
    # oblique ray onto the table top: z drops 1.52 at 0.96 per unit length (lands at y = -0.757, inside the top face)
    assert d[5] == pytest.approx(1.52 / 0.96, abs=1e-12) and body[gb[g[5]]] == "table"


def test_camera_sky_colour_anchor():
    """Weak colour anchor (docs/getting_started.ipynb:414-416): the top rows of the wrist d405 RGB frame show the sky,
    `[169, 224, 255]` in the reference (gradient skybox 0.44 0.80 1.00 -> white, plus haze).  The restated
    renderer draws the gradient without haze: blue saturated, R < G < B, within 60 counts of the printed value."""
    import os
    from oracle.oracle import OracleModel
    from stretch_mujoco_b200 import blob
    raw = blob.read_bytes(os.path.join(os.path.dirname(__file__), "golden", "stretch_default_scene_render.ssm.z"))
    A, names = blob.unpack(raw)
    om = OracleModel(raw)
    om.set_options(enable_lidar=False)
    q = A["qpos0"][None].copy(); v = np.zeros((1, om.nv)); w = np.zeros((1, om.nv))
    o = om.step(q, v, A["key_ctrl"][0][None].copy(), w, nsteps=1600, want=("xpos", "xquat"))       # t = 3.2 s
    cam = names[compiler.OBJ_CAMERA].index("d405_rgb")
    rgb, depth = om.render(o["xpos"], o["xquat"], cam, 96, 54, 58.0)
    top = rgb[0, 0].astype(float).mean(0)
    assert top[2] > 240 and top[0] < top[1] < top[2]
    assert np.abs(top - np.array([169, 224, 255])).max() < 60, top


def test_startup_base_drift_anchor(blob_default_scene):
    """The only contact-sensitive number the reference printed: after `home` from qpos0 in the default scene the base
    has been kicked to x = -0.0122, y = 0.0044, theta = -0.0650 by t = 8.26 s (docs/getting_started.ipynb:729-751) --
    the stowed wrist interpenetrates the base hull by 5 cm at qpos0 and is pushed out through MPR + multiccd contacts
    in the first ~50 steps.  That transient is chaotic (profiles/parity_r2.md: two fp64 oracle runs that differ by
    fp32 rounding separate within tens of steps), so the anchor pins the SIZE and DIRECTION of the kick, not its
    digits: the oracle gives x = -0.0061, y = +0.0054, theta = +0.020 (before multiccd / multi-point plane-mesh
    contacts existed in the oracle the base did not move at all)."""
    from oracle.oracle import OracleModel
    from stretch_mujoco_b200 import blob
    A, _ = blob.unpack(blob_default_scene)
    om = OracleModel(blob_default_scene)
    om.set_options(enable_lidar=False)
    q = A["qpos0"][None].copy(); v = np.zeros((1, om.nv)); w = np.zeros((1, om.nv)); t = np.zeros(1)
    om.step(q, v, A["key_ctrl"][0][None].copy(), w, t, nsteps=4130)
    x, y = q[0, 0], q[0, 1]
    theta = np.arctan2(2 * (q[0, 3] * q[0, 6] + q[0, 4] * q[0, 5]), 1 - 2 * (q[0, 5] ** 2 + q[0, 6] ** 2))
    assert -0.0122 * 3 < x < -0.0122 / 3                  # kicked backwards by about a centimetre
    assert 0.0044 / 3 < y < 0.0044 * 3                    # and a few millimetres to the left
    assert abs(theta) < 0.0650 * 1.5                      # yaw of the same order (its sign is not reproduced)
    assert np.abs(v[0, :6]).max() < 1e-6                  # and the base is at rest again (reference: ~1e-8)


def _table_view(A, names, om):
    """Head camera panned to the robot's right and tilted down: the wood table of scene.xml fills the lower half."""
    jn = names[compiler.OBJ_JOINT]
    q = A["qpos0"][None].copy()
    q[0, A["jnt_qposadr"][jn.index("joint_head_pan")]] = -1.57
    q[0, A["jnt_qposadr"][jn.index("joint_head_tilt")]] = -0.6
    o = om.forward(q, np.zeros((1, om.nv)), A["key_ctrl"][0][None].copy(), None, want=("xpos", "xquat"))
    return o, names[compiler.OBJ_CAMERA].index("d435i_camera_rgb")


def test_wood_table_texture():
    """`<texture type="2d" file="wood.png"/>` behind the table's material (models/scene.xml:15-16,25): the table top carries
    the image (planar x-y projection of the box frame, bilinear, GL_REPEAT) modulated by the lighting.  Against the same model
    with the texture switched off: only table pixels change, they show the grain, and their hue is the one the reference
    printed for the bottom rows of its d405 frame, [160, 133, 100] (docs/getting_started.ipynb:414-416: R/B 1.6, G/B 1.33)."""
    import os
    from oracle.oracle import OracleModel
    from stretch_mujoco_b200 import blob
    raw = blob.read_bytes(os.path.join(os.path.dirname(__file__), "golden", "stretch_default_scene_render.ssm.z"))
    A, names = blob.unpack(raw)
    box = (A["geom_tex"][:, 0] >= 0) & (A["geom_tex"][:, 3] != 2)           # the table; the robot's stickers use their UV sets (mode 2)
    assert int(box.sum()) == 1
    t = int(A["geom_tex"][box, 0][0])
    assert A["tex_w"][t] == 512 and A["tex_h"][t] == 512
    om = OracleModel(raw); om.set_options(enable_lidar=False)
    o, cam = _table_view(A, names, om)
    rgb, depth = om.render(o["xpos"], o["xquat"], cam, 96, 72, 42.0)
    A2 = dict(A); A2["geom_tex"] = A["geom_tex"].copy(); A2["geom_tex"][box, 0] = -1
    om2 = OracleModel(blob.pack(A2, names)); om2.set_options(enable_lidar=False)
    rgb2, depth2 = om2.render(o["xpos"], o["xquat"], cam, 96, 72, 42.0)
    assert np.array_equal(depth, depth2)
    changed = np.abs(rgb.astype(int) - rgb2.astype(int)).max(-1)[0] > 0
    assert 0.3 < changed.mean() < 0.8                                    # the table top, nothing else
    assert 0.5 < depth[0][changed].min() and depth[0][changed].max() < 2.0
    tex = rgb[0][changed].astype(float)
    mean = tex.mean(0)
    assert 1.4 < mean[0] / mean[2] < 1.8 and 1.2 < mean[1] / mean[2] < 1.45, mean
    assert tex.std(0).min() > 5.0                                        # wood grain, not a flat colour
    assert tex.std(0).min() > 1.5 * rgb2[0][changed].std(0).max()        # ... against the smooth shading of the untextured table


def test_aruco_stickers_are_textured():
    """The robot's ArUco / label stickers are small meshes with UV sets behind `<material texture=...>` (models/stretch.xml:82-126,
    405-429): the wrist camera looks at the two finger markers.  Against the same model with textures off, the pixels that change
    are the stickers, and they show the marker's black cells and white border instead of a flat colour."""
    import os
    from oracle.oracle import OracleModel
    from stretch_mujoco_b200 import blob
    raw = blob.read_bytes(os.path.join(os.path.dirname(__file__), "golden", "stretch_empty_floor_render.ssm.z"))
    A, names = blob.unpack(raw)
    assert int((A["geom_tex"][:, 3] == 2).sum()) >= 8 and A["rmesh_uv"].shape == (len(A["rmesh_face"]), 6)
    om = OracleModel(raw); om.set_options(enable_lidar=False)
    q = A["qpos0"][None].copy()
    o = om.step(q, np.zeros((1, om.nv)), A["key_ctrl"][0][None].copy(), np.zeros((1, om.nv)), nsteps=800, want=("xpos", "xquat"))
    A2 = dict(A); A2["geom_tex"] = A["geom_tex"].copy(); A2["geom_tex"][:, 0] = -1
    om2 = OracleModel(blob.pack(A2, names)); om2.set_options(enable_lidar=False)
    cam = names[compiler.OBJ_CAMERA].index("d405_rgb")
    rgb, _ = om.render(o["xpos"], o["xquat"], cam, 120, 68, 58.0)
    rgb2, _ = om2.render(o["xpos"], o["xquat"], cam, 120, 68, 58.0)
    changed = np.abs(rgb.astype(int) - rgb2.astype(int)).max(-1)[0] > 8
    px = rgb[0][changed]
    assert 30 <= changed.sum() <= 400                                   # two finger markers, a few dozen pixels at this size
    assert (px.max(-1) < 60).sum() >= 20 and (px.min(-1) > 150).sum() >= 4   # black cells and white border
