"""Scene loader / model compiler checks.  CPU only.
Tests that read the reference's MJCF files skip when /root/reference is not mounted."""
import os

import numpy as np
import pytest

from conftest import have_reference
from stretch_mujoco_b200 import blob, compiler
from stretch_mujoco_b200.mjcf import quat2mat

needs_ref = pytest.mark.skipif(not have_reference(), reason="reference MJCF files not available")


@pytest.fixture(scope="module")
def model_E():
    from stretch_mujoco_b200 import scenes
    return scenes.compile_empty_floor(with_render=False)


def test_blob_roundtrip(arrays_E, blob_empty_floor):
    A, names = arrays_E
    again = blob.pack(A, names)
    B, names2 = blob.unpack(again)
    assert names == names2 and A.keys() == B.keys()
    for k in A:
        assert np.array_equal(A[k], B[k]), k


def test_golden_sizes(arrays_E):
    A, _ = arrays_E
    nq, nv, nu, nbody, njnt, ngeom, nsite, ncam, nten, neq, nsens, nsd, nkey, nM, npair, nmesh = A["sizes"]
    # nq 27, nv 26, nu 10, nbody 38, njnt 21, nM 260; 360 lidar sites + IMU; 848+51 pairs
    assert (nq, nv, nu, nbody, njnt, nM) == (27, 26, 10, 38, 21, 260)
    assert (nsite, nsens, nsd, ncam, neq, nkey) == (361, 362, 366, 5, 5, 2)
    assert npair == 899 and ngeom == 128


def test_default_scene_sizes(blob_default_scene):
    A, _ = blob.unpack(blob_default_scene)
    assert tuple(A["sizes"][:6]) == (48, 44, 10, 43, 24, 154)  # SURVEY.md A.1 default scene.xml


@needs_ref
def test_golden_matches_fresh_compile(model_E, arrays_E):
    A, _ = arrays_E
    for k, v in model_E.arrays.items():
        assert k in A, k
        assert np.allclose(A[k], v, rtol=0, atol=1e-12), k


@needs_ref
def test_masses(model_E):
    m = model_E
    bid = lambda n: m.name2id(compiler.OBJ_BODY, n)
    exp = {"base_link": 25.0, "link_mast": 1.8285, "link_head": 0.833027236718691, "laser": 0.24, "link_lift": 3.0,
           "link_right_wheel": 0.15, "link_arm_l4": 0.168095, "link_arm_l0": 0.28501, "link_wrist_yaw": 0.1445,
           "link_SG3_gripper_body": 0.29, "link_d405": 0.29 + 2 * 3.6e-6, "link_gripper_slider": 0.05,
           "link_gripper_finger_left": 0.10, "link_gripper_finger_right": 0.15, "rubber_tip_left": 0.005,
           "link_head_tilt": 0.262217, "link_SE3_head_nav_cam": 0.028346169831483}
    for n, mass in exp.items():
        assert m.body_mass[bid(n)] == pytest.approx(mass, rel=1e-9), n
    assert m.body_mass.sum() == pytest.approx(33.695, abs=1e-3)


@needs_ref
def test_actuators_and_limits(model_E):
    m = model_E
    names = m.names[compiler.OBJ_ACTUATOR]
    assert names == ["left_wheel_vel", "right_wheel_vel", "lift", "arm", "wrist_yaw", "wrist_pitch", "wrist_roll",
                     "gripper", "head_pan", "head_tilt"]
    a = names.index
    # velocity servo kv 20 with gear 3; position servos inherit the class-level general/position parameters
    assert m.actuator_gainprm[a("left_wheel_vel"), 0] == 20 and m.actuator_biasprm[a("left_wheel_vel"), 2] == -20
    assert m.actuator_gear[a("left_wheel_vel")] == 3
    assert tuple(m.actuator_biasprm[a("lift")]) == (0, -400, -100) and tuple(m.actuator_forcerange[a("lift")]) == (-70, 70)
    assert tuple(m.actuator_biasprm[a("arm")]) == (0, -150, -10) and m.actuator_trntype[a("arm")] == 1
    assert tuple(m.actuator_biasprm[a("gripper")]) == (0, -4000, -124)
    assert tuple(m.actuator_biasprm[a("head_tilt")]) == (0, -10, 0)
    # compiled joint limits quoted in enums/actuators.py:69-89
    j = m.names[compiler.OBJ_JOINT].index
    rng = lambda n: tuple(m.jnt_range[j(n)])
    assert rng("joint_lift") == (0.0, 1.1) and rng("joint_arm_l0") == (0.0, 0.13)
    assert rng("joint_wrist_yaw") == (-1.39, 4.42) and rng("joint_wrist_pitch") == (-1.57, 0.56)
    assert rng("joint_head_pan") == (-4.04, 1.73) and rng("joint_head_tilt") == (-1.53, 0.79)
    assert rng("joint_gripper_slide") == (-0.02, 0.04) and m.jnt_limited[j("joint_left_wheel")] == 0
    assert m.key_ctrl[0].tolist() == [0, 0, 0.6, 0.1, 0, 0, 0, 0, 0, 0]


@needs_ref
def test_frames_at_qpos0(model_E):
    """Frame-convention KATs of SURVEY.md Appendix A.2 (camera / lidar / IMU placement)."""
    m = model_E
    xpos, xquat, _, _ = compiler.fk_numpy(m, m.qpos0)
    cam = m.names[compiler.OBJ_CAMERA].index

    def cam_frame(n):
        c = cam(n); b = m.cam_bodyid[c]
        R = quat2mat(xquat[b]) @ quat2mat(m.cam_quat[c])
        return xpos[b] + quat2mat(xquat[b]) @ m.cam_pos[c], R

    p, R = cam_frame("d435i_camera_rgb")
    assert np.allclose(p, [0.045, -0.003, 1.322], atol=2e-3)
    assert np.allclose(-R[:, 2], [1, 0, 0], atol=2e-2) and np.allclose(R[:, 1], [0, -1, 0], atol=2e-2)
    p, R = cam_frame("d405_rgb")
    assert np.allclose(p, [-0.021, -0.245, 0.176], atol=2e-3)
    assert np.allclose(-R[:, 2], [0.002, -0.984, -0.176], atol=5e-3)
    sid = m.names[compiler.OBJ_SITE].index
    for name, d in (("lidar000", [-1, 0, 0]), ("lidar090", [0, -1, 0]), ("lidar180", [1, 0, 0])):
        s = sid(name); b = m.site_bodyid[s]
        z = (quat2mat(xquat[b]) @ quat2mat(m.site_quat[s]))[:, 2]
        assert np.allclose(z, d, atol=2e-3), name
    s = sid("base_imu"); b = m.site_bodyid[s]
    R = quat2mat(xquat[b]) @ quat2mat(m.site_quat[s])
    assert np.allclose(R[:, 0], [0, -1, 0], atol=1e-3) and np.allclose(R[:, 2], [0, 0, -1], atol=1e-3)


@needs_ref
def test_lidar_sensor_names(model_E):
    names = model_E.names[compiler.OBJ_SENSOR]
    # enums/stretch_sensors.py:33-41 expects base_lidar000 .. base_lidar359
    assert names[0] == "base_gyro" and names[1] == "base_accel"
    assert names[2:] == [f"base_lidar{i:03d}" for i in range(360)]


def test_kitchen_proxy_golden_layout():
    """BASELINE config 4 stand-in (scenes.KITCHEN_PROXY_XML): robot + static box fixtures + one free box, lidar
    ring re-spun at 1000 rays through the loader's replicate override."""
    import os
    from conftest import GOLDEN
    from stretch_mujoco_b200 import blob
    A, names = blob.unpack(blob.read_bytes(os.path.join(GOLDEN, "stretch_kitchen_proxy_render.ssm.z")))
    assert len(A["qpos0"]) == 34 and len(A["dof_bodyid"]) == 32                 # 27 + 7, 26 + 6
    snames = names[3]
    assert snames.count("lidar0000") == 1 and "lidar0999" in snames and "lidar1000" not in snames
    sid = [snames.index(f"lidar{i:04d}") for i in (0, 250, 500)]
    from stretch_mujoco_b200.mjcf import quat2mat
    z = [quat2mat(A["site_quat"][s])[:, 2] for s in sid]                        # ray direction = site +z
    assert np.allclose(z[0], [1, 0, 0], atol=1e-9) and np.allclose(z[1], [0, 1, 0], atol=1e-6) and np.allclose(z[2], [-1, 0, 0], atol=1e-6)
    gnames = names[2]
    for g in ("wall", "wall_left", "wall_right", "counter_main", "stove", "counter_right", "floor"):
        assert g in gnames
    st = np.asarray(A["sensor_type"])
    assert (st == 2).sum() == 1000 and len(A["raygeom_id"]) > 80


@needs_ref
def test_kitchen_proxy_golden_matches_fresh_compile():
    import os
    from conftest import GOLDEN
    from stretch_mujoco_b200 import blob, scenes
    A, _ = blob.unpack(blob.read_bytes(os.path.join(GOLDEN, "stretch_kitchen_proxy_render.ssm.z")))
    m = scenes.compile_kitchen_proxy(with_render=True, lidar_rays=1000)
    for k in ("qpos0", "body_mass", "geom_size", "site_quat", "pair_geom1", "pair_geom2", "rmesh_face"):
        assert np.array_equal(np.asarray(m.arrays[k]), A[k]), k


def test_mjmodel_npz_import_round_trip(tmp_path, blob_empty_floor):
    """a dump of a real mjModel (flat .npz of named arrays) loads through the same path.  Without a
    MuJoCo install the hook is exercised with a dump written from this compiler's own arrays in mjModel's layout
    (mesh graphs, 10-wide actuator parameter rows, 6-vector gear, 2-column trnid)."""
    from stretch_mujoco_b200 import blob as blobmod
    A, names = blobmod.unpack(blob_empty_floor)
    Z = {f: A[f] for f in compiler._MJMODEL_FIELDS if f in A}
    nu = len(A["actuator_trnid"])
    Z["actuator_trnid"] = np.stack([A["actuator_trnid"], -np.ones(nu, np.int32)], 1)
    Z["actuator_gear"] = np.concatenate([A["actuator_gear"][:, None], np.zeros((nu, 5))], 1)
    for k in ("actuator_gainprm", "actuator_biasprm"):
        Z[k] = np.concatenate([A[k], np.zeros((nu, 7))], 1)
    for k in ("opt_timestep", "opt_impratio", "opt_tolerance", "opt_ls_tolerance", "stat_meaninertia", "stat_extent",
              "opt_iterations", "opt_ls_iterations", "opt_cone", "opt_solver", "opt_multiccd"):
        Z[k] = A[k][0]
    Z["opt_gravity"] = A["opt_gravity"]
    Z["sensor_enum"] = np.array([compiler.SENS_GYRO, compiler.SENS_ACCEL, compiler.SENS_RANGE])
    # mesh graphs in mjModel's layout from the hull tables
    mesh_vert, vertadr, graph, graphadr = [], [], [], []
    for k in range(len(A["mesh_hulladr"])):
        vertadr.append(sum(len(v) for v in mesh_vert))
        ha, hn = int(A["mesh_hulladr"][k]), int(A["mesh_hullnum"][k])
        if ha < 0:
            graphadr.append(-1); mesh_vert.append(np.zeros((1, 3)))
            continue
        mesh_vert.append(A["hull_vert"][ha:ha + hn])
        ea = A["hull_edgeadr"][ha:ha + hn + 1]
        edge_local, vert_edgeadr = [], []
        for v in range(hn):
            vert_edgeadr.append(len(edge_local))
            edge_local += list(A["hull_edge"][ea[v]:ea[v + 1]]) + [-1]
        nf = (len(edge_local) - hn + 2) // 3 + 1
        edge_local += [-1] * (hn + 3 * nf - len(edge_local))
        graphadr.append(len(graph))
        graph += [hn, nf] + vert_edgeadr + list(range(hn)) + edge_local + [0] * (3 * nf)
    Z.update(mesh_vert=np.concatenate(mesh_vert), mesh_vertadr=np.array(vertadr), mesh_vertnum=np.array([len(v) for v in mesh_vert]),
             mesh_graph=np.array(graph), mesh_graphadr=np.array(graphadr))
    for objtype, prefix in compiler._OBJ_NAMES.items():
        Z["names_" + prefix] = np.array(names.get(objtype, []))
    path = str(tmp_path / "mjmodel.npz")
    np.savez_compressed(path, **Z)
    m2 = compiler.load_mjmodel_npz(path)
    ours = compiler.Model(); ours.arrays, ours.names = A, names
    diff = compiler.compare_models(ours, m2)
    assert diff == {}, diff
    assert np.array_equal(m2.arrays["pair_geom1"], A["pair_geom1"]) and np.array_equal(m2.arrays["hull_edge"], A["hull_edge"])
    assert np.array_equal(m2.arrays["sizes"], A["sizes"])
    assert m2.name2id(compiler.OBJ_BODY, "base_link") == ours.name2id(compiler.OBJ_BODY, "base_link")
    # and the imported model steps in the oracle exactly like the compiled one
    from oracle.oracle import OracleModel
    qs = []
    for arrays, nm in ((A, names), (m2.arrays, m2.names)):
        om = OracleModel(blobmod.pack(arrays, nm)); om.set_options(enable_lidar=False)
        q = arrays["qpos0"][None].copy(); v = np.zeros((1, om.nv)); w = np.zeros((1, om.nv))
        om.step(q, v, arrays["key_ctrl"][0][None].copy(), w, nsteps=50)
        qs.append(q)
    assert np.array_equal(qs[0], qs[1])


@needs_ref
def test_stretch_mj_3_3_0_variant_compiles_to_the_same_robot():
    """models/stretch_mj_3.3.0.xml moves the shell-inertia switch from the geoms to the mesh assets
    (inertia="shell", :129-222), gives the cameras their own fovy (:378,456,463) and drops the lidar replicate: same
    bodies, dofs and mass properties as stretch.xml, five cameras with fovy 58 / 69 where the file sets them."""
    from stretch_mujoco_b200 import compiler, mjcf, scenes
    d = scenes.models_dir()
    new = compiler.compile_scene(mjcf.Scene.from_xml_path(os.path.join(d, "stretch_mj_3.3.0.xml")), with_render=False)
    old = compiler.compile_scene(mjcf.Scene.from_xml_path(os.path.join(d, "stretch.xml")), with_render=False)
    a, b = new.arrays, old.arrays
    assert list(a["sizes"][:6]) == list(b["sizes"][:6])                       # nq nv nu nbody njnt ngeom
    for k in ("body_mass", "body_inertia", "body_ipos", "body_iquat", "jnt_range", "actuator_biasprm"):
        assert np.array_equal(a[k], b[k]), k
    assert a["sizes"][6] == 1 and b["sizes"][6] == 361                        # sites: the lidar replicate is gone
    cams = new.names[4]
    fovy = dict(zip(cams, a["cam_fovy"]))
    assert fovy["d405_rgb"] == 58.0 and fovy["d435i_camera_rgb"] == 58.0 and fovy["nav_camera_rgb"] == 69.0
