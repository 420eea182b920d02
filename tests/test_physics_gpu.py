"""Parity of the CUDA physics path (through the C ABI) against the fp64 CPU oracle.  GPU only.

Tolerances (fp32 device vs fp64 oracle, stated per test):
  * single forward pass: mass matrix 1e-5 rel, contact pair list bit-exact, contact distance 1e-6 m,
    joint-space constraint force 1e-3 rel;
  * 1000-step rollouts from the settled `home` pose with per-env random targets: envs whose whole
    rollout has only plane contacts (wheels, caster) must agree to 1e-4 relative in qpos
    (north-star bar); envs that enter convex mesh contact (single-point MPR contacts are
    discontinuous in the configuration) are bounded statistically.
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gpu(blob_empty_floor):
    from stretch_mujoco_b200 import engine
    return engine.DeviceModel(blob_empty_floor, 0)


def _targets(rng, home, nenv):
    ctrl = np.tile(home, (nenv, 1))
    ctrl[:, 0:2] = rng.uniform(-3, 3, (nenv, 2)); ctrl[:, 2] = rng.uniform(0.3, 1.0, nenv); ctrl[:, 3] = rng.uniform(0, 0.4, nenv)
    ctrl[:, 4] = rng.uniform(-1, 3, nenv); ctrl[:, 5] = rng.uniform(-1, 0.5, nenv); ctrl[:, 6] = rng.uniform(-2, 2, nenv)
    ctrl[:, 7] = rng.uniform(-0.02, 0.04, nenv); ctrl[:, 8] = rng.uniform(-3, 1.5, nenv); ctrl[:, 9] = rng.uniform(-1.4, 0.7, nenv)
    ctrl[0] = home
    return ctrl


def _load(B, qpos, qvel, warm, ctrl):
    B.qpos.copy_(torch.tensor(qpos, dtype=torch.float32)); B.qvel.copy_(torch.tensor(qvel, dtype=torch.float32))
    B.qacc_warmstart.copy_(torch.tensor(warm, dtype=torch.float32)); B.ctrl.copy_(torch.tensor(ctrl, dtype=torch.float32))
    B.time.zero_(); B.env_flags.zero_()
    f = lambda t: t.cpu().numpy().astype(np.float64)
    return f(B.qpos), f(B.qvel), f(B.qacc_warmstart), f(B.ctrl)  # fp32-rounded copies for the oracle


def test_forward_matches_oracle_at_qpos0(gpu, oracle_E, arrays_E):
    from stretch_mujoco_b200 import engine
    A, _ = arrays_E
    nenv = 32
    rng = np.random.default_rng(0)
    B = engine.Batch(gpu, nenv, debug=True)
    lo, hi = A["actuator_ctrlrange"][:, 0], A["actuator_ctrlrange"][:, 1]
    ctrl = rng.uniform(lo, hi, size=(nenv, gpu.nu))
    qpos, qvel, warm, ctrl = _load(B, np.tile(A["qpos0"], (nenv, 1)), rng.normal(scale=0.05, size=(nenv, gpu.nv)),
                                   np.zeros((nenv, gpu.nv)), ctrl)
    B.forward(); torch.cuda.synchronize()
    o = oracle_E.forward(qpos, qvel, ctrl, warm, maxcon=B.maxcon,
                         want=("M", "qacc_smooth", "ncon", "nefc", "contact_geom", "contact_dist", "qfrc_constraint", "qacc"))
    M = B.dbg["M"].cpu().numpy()
    assert np.abs(M - o["M"]).max() <= 1e-6 * np.abs(o["M"]).max()
    dM = np.sqrt(np.einsum("eii->ei", o["M"]))
    assert (np.abs(M - o["M"]) / (dM[:, :, None] * dM[:, None, :])).max() < 5e-6      # every entry, relative to sqrt(Mii Mjj): light links included
    assert np.array_equal(B.ncon.cpu().numpy(), o["ncon"])
    assert np.array_equal(B.contact_geom.cpu().numpy(), o["contact_geom"])          # bit-exact pair indexing
    assert np.array_equal(B.dbg["nefc"].cpu().numpy(), o["nefc"])
    # the first contact of every geom pair is the plain narrowphase result: 1e-6 m.  The following contacts of a
    # convex pair are multiccd contacts (stretch.xml:8): MPR queries of poses perturbed by 1e-3 rad, whose portal
    # (hence depth) reacts to the fp32 rounding of the inputs -- they are bounded at 0.5 mm.
    cg = o["contact_geom"]; dd = np.abs(B.contact_dist.cpu().numpy() - o["contact_dist"])
    first = np.ones(cg.shape[:2], bool); first[:, 1:] = np.any(cg[:, 1:] != cg[:, :-1], axis=2)
    plane = A["geom_type"][np.maximum(cg[:, :, 0], 0)] == 0
    assert dd[first | plane].max() < 1e-6 and dd.max() < 5e-4
    qs = B.dbg["qacc_smooth"].cpu().numpy()
    e = qs - o["qacc_smooth"]
    en = np.sqrt(np.einsum("ei,eij,ej->e", e, o["M"], e) / np.einsum("ei,eij,ej->e", o["qacc_smooth"], o["M"], o["qacc_smooth"]))
    assert en.max() < 5e-6, en.max()                  # energy norm (measured 4e-7 with the COM-form mass matrix; 8e-5 before it)
    assert np.abs(e).max() <= 1e-4 * np.abs(o["qacc_smooth"]).max(), np.abs(e).max() / np.abs(o["qacc_smooth"]).max()
    fc = B.dbg["qfrc_constraint"].cpu().numpy()
    # qpos0 has the stowed wrist inside the base hull: the multiccd contacts of that pair carry up to 0.5 mm of MPR portal noise in
    # their depth (above), which the stiff contact turns into force -- measured 4.3e-4 of the largest constraint force in every env
    rel = np.abs(fc - o["qfrc_constraint"]).max(axis=1) / np.abs(o["qfrc_constraint"]).max(axis=1)
    assert np.median(rel) < 1e-3 and rel.max() < 2e-3, (np.median(rel), rel.max())


def test_forward_matches_oracle_from_home(gpu, oracle_E, arrays_E, settled_home_E):
    from stretch_mujoco_b200 import engine
    A, _ = arrays_E
    nenv = 64
    rng = np.random.default_rng(1)
    q0, v0, w0, home = settled_home_E
    B = engine.Batch(gpu, nenv, debug=True)
    qpos = np.tile(q0, (nenv, 1)); qpos[:, 7:] += rng.normal(scale=0.01, size=(nenv, gpu.nq - 7))
    qpos, qvel, warm, ctrl = _load(B, qpos, np.tile(v0, (nenv, 1)) + rng.normal(scale=0.02, size=(nenv, gpu.nv)),
                                   np.tile(w0, (nenv, 1)), _targets(rng, home, nenv))
    B.forward(); torch.cuda.synchronize()
    o = oracle_E.forward(qpos, qvel, ctrl, warm, maxcon=B.maxcon, want=("ncon", "contact_geom", "qfrc_constraint", "qacc", "nefc"))
    assert np.array_equal(B.contact_geom.cpu().numpy(), o["contact_geom"])
    assert np.array_equal(B.dbg["nefc"].cpu().numpy(), o["nefc"])
    fc = B.dbg["qfrc_constraint"].cpu().numpy()
    rel = np.abs(fc - o["qfrc_constraint"]).max(axis=1) / np.abs(o["qfrc_constraint"]).max(axis=1)
    assert np.median(rel) < 1e-4 and rel.max() < 5e-3
    # one full step (implicitfast) from the same states: velocities agree to 2e-4 rad/s (fp32 noise on the
    # 8e-7 kg m^2 rubber-tip joints dominates), positions to 1e-6
    B.step(1); torch.cuda.synchronize()
    oracle_E.step(qpos, qvel, ctrl, warm, nsteps=1)
    dv = np.abs(B.qvel.cpu().numpy() - qvel).max(axis=1)
    assert np.median(dv) < 1e-4 and dv.max() < 5e-4
    assert np.abs(B.qpos.cpu().numpy() - qpos).max() < 2e-6


def test_forward_parity_on_random_rollout_states(gpu, oracle_E, arrays_E):
    """bench.py's workload (uniform-random ctrl, redrawn every 50 steps) drives the robot into joint
    limits, self-contact and tipping.  After 207 steps every env's state is handed to the oracle and ONE
    forward pass is compared: contact pair lists bit-exact; qacc of the fp32 Newton solve within 5e-5
    (median) / 1e-3 (99th percentile; multiccd contacts of perturbed poses carry the MPR portal noise) of the fp64 solve relative to the env's largest acceleration.
    The remaining outliers must all be envs where a deeply overlapping convex pair (gripper linkage
    hulls) has its MPR portal land on a different face in fp32 than in fp64 -- a discontinuity of the
    single-point MPR contact, not a solver error (found with tests/compare_smooth.py)."""
    import bench
    from stretch_mujoco_b200 import engine
    A, _ = arrays_E
    nenv = 1024
    B = engine.Batch(gpu, nenv, debug=True)
    dev = B.qpos.device
    lo = torch.tensor(A["actuator_ctrlrange"][:, 0], dtype=torch.float64, device=dev)
    hi = torch.tensor(A["actuator_ctrlrange"][:, 1], dtype=torch.float64, device=dev)
    for p in range(4):
        B.ctrl.copy_(bench.ctrl_torch(0, 0, nenv, p, lo, hi, dev)); B.step(50)
    B.step(7)
    torch.cuda.synchronize()
    f = lambda t: t.cpu().numpy().astype(np.float64)
    q, v, w, c = f(B.qpos), f(B.qvel), f(B.qacc_warmstart), f(B.ctrl)
    B.forward(); torch.cuda.synchronize()
    oracle_E.set_options(enable_lidar=False)
    o = oracle_E.forward(q, v, c, w, maxcon=B.maxcon, want=("qacc", "ncon", "contact_geom", "contact_dist", "contact_frame", "nefc", "flags"))
    ok = (o["flags"] & 2) == 0                     # envs inside the contact capacity on the oracle side
    assert ok.mean() > 0.99
    assert np.array_equal(B.ncon.cpu().numpy()[ok], o["ncon"][ok])
    assert np.array_equal(B.contact_geom.cpu().numpy()[ok], o["contact_geom"][ok])      # bit-exact pair indexing
    assert np.array_equal(B.dbg["nefc"].cpu().numpy()[ok], o["nefc"][ok])
    qa = B.qacc.cpu().numpy()
    err = np.abs(qa - o["qacc"]).max(1) / (np.abs(o["qacc"]).max(1) + 1e-3)
    err = np.where(ok, err, 0.0)
    assert np.median(err) < 5e-5 and np.quantile(err, 0.99) < 1e-3, (np.median(err), np.quantile(err, 0.99))   # measured 2.2e-5 / 1.3e-4
    gn = B.dbg["contact_normal"].cpu().numpy()
    live = np.arange(B.maxcon)[None, :] < o["ncon"][:, None]
    ndiff = np.where(live, np.abs(gn - o["contact_frame"]).max(2), 0.0).max(1)      # largest normal mismatch per env
    outliers = err > 1e-2
    assert outliers.mean() < 0.01
    assert np.all(ndiff[outliers] > 1e-2), "a qacc outlier without an MPR face flip would be a solver error"
    assert np.median(ndiff) < 1e-5


def test_rollout_1000_steps_from_home(gpu, oracle_E, arrays_E, settled_home_E):
    """64 envs x 1000 steps, per-env random arm/wrist/head targets and gentle base motion.
    Contact activation is a discontinuity of the dynamics (a wheel that touches at -1e-9 m in
    fp64 and misses at +1e-9 m in fp32 changes the next step), so the 1e-4 bar is asserted for
    every env whose contact list agreed at all 100 checkpoints, and those must be the bulk."""
    from stretch_mujoco_b200 import engine
    A, _ = arrays_E
    nenv = 64
    rng = np.random.default_rng(0)
    q0, v0, w0, home = settled_home_E
    B = engine.Batch(gpu, nenv)
    tg = _targets(rng, home, nenv)
    tg[:, 0:2] = rng.uniform(-0.5, 0.5, (nenv, 2)); tg[:, 7] = rng.uniform(0.0, 0.04, nenv)
    qpos, qvel, warm, ctrl = _load(B, np.tile(q0, (nenv, 1)), np.tile(v0, (nenv, 1)), np.tile(w0, (nenv, 1)), tg)
    t = np.zeros(nenv)
    pairs_equal = np.ones(nenv, bool)
    for k in range(100):
        B.step(10); torch.cuda.synchronize()
        o = oracle_E.step(qpos, qvel, ctrl, warm, t, nsteps=10, maxcon=B.maxcon, want=("contact_geom", "flags"))
        pairs_equal &= np.all(B.contact_geom.cpu().numpy() == o["contact_geom"], axis=(1, 2))
    err = np.abs(B.qpos.cpu().numpy() - qpos)
    rel = (err / np.maximum(np.abs(qpos), 1.0)).max(axis=1)
    print("rollout parity: pairs always equal in %d/%d envs; rel qpos err median %.2e, max over matching envs %.2e, max %.2e"
          % (pairs_equal.sum(), nenv, np.median(rel), rel[pairs_equal].max(), rel.max()))
    assert pairs_equal.sum() >= int(0.95 * nenv)                  # measured: 64/64
    assert rel[pairs_equal].max() < 1e-4, f"max rel qpos error {rel[pairs_equal].max():.2e}"   # measured 3.1e-5 (north-star bar: 1e-4)
    assert np.median(rel) < 2e-5 and rel.max() < 5e-3            # measured: median 7.0e-6, max 3.1e-5
    assert float(B.time[0]) == pytest.approx(2.0, abs=1e-4)
    assert int(B.env_flags.max()) == 0 and int(o["flags"].max()) == 0


def test_bad_state_guard_resets_env(gpu, arrays_E):
    from stretch_mujoco_b200 import engine
    A, _ = arrays_E
    B = engine.Batch(gpu, 4)
    B.qvel[2, 8] = float("nan")
    B.qpos[3, 0] = 1e12
    B.step(1); torch.cuda.synchronize()
    flags = B.env_flags.cpu().numpy()
    assert flags[0] == 0 and flags[1] == 0 and flags[2] & 1 and flags[3] & 1
    assert torch.isfinite(B.qpos).all() and torch.isfinite(B.qvel).all()
    assert np.abs(B.qpos[3].cpu().numpy() - A["qpos0"]).max() < 0.1   # one step from qpos0 after the reset


def test_reset_and_keyframe(gpu, arrays_E):
    from stretch_mujoco_b200 import engine
    A, _ = arrays_E
    B = engine.Batch(gpu, 8)
    B.step(10)
    mask = torch.zeros(8, dtype=torch.int32, device="cuda"); mask[1] = 1; mask[5] = 1
    B.reset(mask, key=1)  # stow
    torch.cuda.synchronize()
    assert np.allclose(B.ctrl[1].cpu().numpy(), A["key_ctrl"][1]) and np.allclose(B.ctrl[0].cpu().numpy(), 0)
    assert np.allclose(B.qpos[5].cpu().numpy(), A["qpos0"], atol=1e-6) and float(B.time[5]) == 0.0 and float(B.time[0]) > 0


def test_determinism(gpu):
    from stretch_mujoco_b200 import engine
    outs = []
    for _ in range(2):
        B = engine.Batch(gpu, 16)
        B.reset(key=0); B.step(200); torch.cuda.synchronize()
        outs.append(B.qpos.clone())
    assert torch.equal(outs[0], outs[1])


def test_status_and_commands_follow_reference_semantics(gpu, arrays_E):
    """P1/P2 rows: pull_status / push_command / BaseController (mujoco_server.py:93-176,465-578)."""
    from stretch_mujoco_b200.simulator import StretchMujocoSimulator
    A, _ = arrays_E
    sim = StretchMujocoSimulator(model=gpu, nenv=4)
    with pytest.raises(ConnectionError):
        sim.pull_status()
    sim.start()
    with pytest.raises(Exception):
        sim.move_to("base_translate", 0.1)
    with pytest.raises(Exception):
        sim.move_by("left_wheel_vel", 0.1)
    sim.step(1500)
    s = sim.pull_status()
    assert float(s.time[0]) == pytest.approx(3.0, abs=1e-3)
    assert float(s.lift.pos[0]) == pytest.approx(0.589, abs=3e-3) and float(s.arm.pos[0]) == pytest.approx(0.1, abs=1e-3)
    assert float(s.gripper.pos[0]) == pytest.approx(-0.064, abs=2e-3)   # sim 0 -> real range (config.py:4-5)
    # move_to with per-env targets, move_by relative to the current length, gripper in the real range
    sim.move_to("lift", torch.tensor([0.4, 0.5, 0.7, 0.8]))
    sim.move_by("arm", 0.1)
    sim.move_to("gripper", 0.56, env_ids=torch.tensor([1]))
    sim.step(1)
    c = sim.batch.ctrl.cpu().numpy()
    assert np.allclose(c[:, 2], [0.4, 0.5, 0.7, 0.8]) and np.allclose(c[:, 3], 0.2, atol=2e-3)
    assert c[1, 7] == pytest.approx(0.04, abs=1e-6) and c[0, 7] == pytest.approx(0.0, abs=1e-6)
    cmd = sim.batch.command
    assert float(cmd[:, 0:10].abs().sum() + cmd[:, 20:30].abs().sum() + cmd[:, 40].abs().sum() + cmd[:, 43].abs().sum()) == 0.0  # triggers consumed
    assert sim.wait_until_at_setpoint("lift", timeout=8.0)
    # set_base_velocity -> wheel ctrl through diff-drive inverse kinematics; status reports it back
    sim.set_base_velocity(0.1, 0.0)
    sim.step(500)
    c = sim.batch.ctrl.cpu().numpy()
    assert np.allclose(c[:, 0:2], 0.1 / 0.0508, atol=1e-4)
    s = sim.pull_status()
    # gear=3 quirk: the servo tracks 3*qdot -> reported speed approaches the command from
    # below; the 35 N m wheel frictionloss (stretch.xml:17) leaves a steady-state gap of ~0.03 m/s
    assert 0.04 < float(s.base.x_vel[0]) <= 0.1
    # base_translate by +0.05 m: closed loop runs at 0.3 m/s until the displacement is reached, then stops
    x0 = s.base.x.clone(); y0 = s.base.y.clone()
    sim.move_by("base_translate", 0.05)
    for _ in range(400):
        sim.step(5)
        if float(sim.batch.base_state[:, 0].abs().sum()) == 0:
            break
    s = sim.pull_status()
    d = torch.sqrt((s.base.x - x0) ** 2 + (s.base.y - y0) ** 2)
    assert float(sim.batch.base_state[:, 0].abs().sum()) == 0 and torch.all(d >= 0.045) and torch.all(d < 0.08)
    assert np.allclose(sim.batch.ctrl[:, 0:2].cpu().numpy(), 0.0)
    sim.stop()
    assert not sim.is_running()


def test_scheduling_does_not_leak_into_results(gpu, arrays_E, monkeypatch):
    """The launch policy (cost-sorted work queue, 2-step launches, env sets on side streams) only decides
    which warp simulates which env and when: a batch stepped with the policy switched off (one 50-step
    launch, identity order, one set) must stay bit-identical to the default one over a random-ctrl rollout."""
    import bench
    from stretch_mujoco_b200 import engine
    A, _ = arrays_E
    nenv = 2048
    B1 = engine.Batch(gpu, nenv)
    for k, v in (("SS_SETS", "1"), ("SS_CHUNK", "50"), ("SS_NOSORT", "1")):
        monkeypatch.setenv(k, v)
    B2 = engine.Batch(gpu, nenv)
    dev = B1.qpos.device
    lo = torch.tensor(A["actuator_ctrlrange"][:, 0], dtype=torch.float64, device=dev)
    hi = torch.tensor(A["actuator_ctrlrange"][:, 1], dtype=torch.float64, device=dev)
    for p in range(5):
        c = bench.ctrl_torch(0, 0, nenv, p, lo, hi, dev)
        B1.ctrl.copy_(c); B2.ctrl.copy_(c)
        B1.step(50); B2.step(50)
        torch.cuda.synchronize()
        assert torch.equal(B1.qpos, B2.qpos) and torch.equal(B1.qvel, B2.qvel) and torch.equal(B1.qacc_warmstart, B2.qacc_warmstart)
        assert torch.equal(B1.xpos, B2.xpos) and torch.equal(B1.contact_geom, B2.contact_geom) and torch.equal(B1.time, B2.time)
    assert B1.launches > B2.launches


def test_batch_size_invariance_at_the_bench_size(gpu, arrays_E):
    """Envs are independent units: env i of BASELINE config 2 (4096 envs, bench.py's ctrl stream from qpos0, two control
    periods) is bit-identical to the same env simulated in a batch of 1, 7, 33 or 1187 envs -- whatever the CTA shapes,
    env sets and cost-sorted schedule of the two batches (a size-independent property checked at the full size)."""
    import bench
    from stretch_mujoco_b200 import engine
    A, _ = arrays_E
    dev = torch.device("cuda")
    lo = torch.tensor(A["actuator_ctrlrange"][:, 0], dtype=torch.float64, device=dev)
    hi = torch.tensor(A["actuator_ctrlrange"][:, 1], dtype=torch.float64, device=dev)

    def rollout(env0, n):
        B = engine.Batch(gpu, n)
        for p in range(2):
            B.ctrl.copy_(bench.ctrl_torch(0, env0, n, p, lo, hi, dev)); B.step(50)
        torch.cuda.synchronize()
        return B.qpos.clone(), B.qvel.clone(), B.qacc_warmstart.clone(), B.contact_geom.clone()

    big = rollout(0, 4096)
    for env0, n in ((0, 1), (5, 7), (100, 33), (2900, 1187)):
        small = rollout(env0, n)
        for a, b in zip(big, small):
            assert torch.equal(a[env0:env0 + n], b), (env0, n)
