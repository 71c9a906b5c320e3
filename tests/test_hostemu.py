"""Kernel SOURCE checked against the oracle without a GPU (CPU only).

tests/hostemu/libmjb_hostemu.so is the product's kernel source (mujoco_b200/csrc/mjb_*.h) compiled
by g++ with the same C ABI.  Built with -ffp-contract=off it performs the same IEEE operations in the
same order as the reference engine, so every mjData field and every trajectory must be BIT-EXACT.
This is a test of the algorithm restatement; GPU parity proper is tests/test_gpu_parity.py."""
import os

import numpy as np
import pytest

import mujoco_b200 as mb
from mjb_util import ANT, HOSTEMU, HUMANOID, ROOT, compare_forward, hostemu_lib, make_pair, perturbed_states
from oracle_util import available

pytestmark = pytest.mark.skipif(not (available() and os.path.exists(HOSTEMU)), reason="oracle or hostemu not built")


@pytest.mark.parametrize("solver", [mb.SOLVER_PGS, mb.SOLVER_NEWTON])
def test_forward_every_field_bit_exact(solver):
    m, b, o = make_pair(HUMANOID, solver, library=hostemu_lib(), nenv=12)
    states = perturbed_states(o, 12, seed=3, height=[0.2, 0.3, 0.45, 0.9, 1.3, 2.0])
    ctrl = np.random.default_rng(5).uniform(-1.5, 1.5, (12, o.size("nu")))   # beyond ctrlrange: clamp path
    compare_forward(b, o, states, ctrl, rtol=0, exact=True, check_dual=(solver == mb.SOLVER_PGS))


@pytest.mark.parametrize("model,tag,solver", [("humanoid", "pgs", mb.SOLVER_PGS), ("humanoid", "newton", mb.SOLVER_NEWTON),
                                              ("ant", "newton", mb.SOLVER_NEWTON)])
def test_golden_trajectory_bit_exact(model, tag, solver):
    g = np.load(os.path.join(ROOT, "tests", "golden", "%s_%s_traj.npz" % (model, tag)))
    m = mb.Model(os.path.join(ROOT, "models", model + ".mjb"), library=hostemu_lib())
    m.set_option("solver", solver)
    b = mb.Batch(m, g["state0"].shape[0])
    out = b.rollout(g["state0"], g["ctrl"])
    assert np.array_equal(out, g["states"])


@pytest.mark.parametrize("solver", [mb.SOLVER_PGS, mb.SOLVER_NEWTON])
def test_contact_rich_rollout_bit_exact(solver):
    nenv, nstep = 16, 120
    m, b, o = make_pair(HUMANOID, solver, library=hostemu_lib(), nenv=nenv)
    s0 = perturbed_states(o, nenv, seed=11, height=[0.2, 0.3, 0.5, 0.8], qvel_std=0.5, qpos_std=0.2)
    ctrl = np.random.default_rng(12).uniform(-1, 1, (nenv, nstep, o.size("nu")))
    out = b.rollout(s0, ctrl)
    ref, stats, _ = o.rollout(s0, ctrl, nthread=4)
    assert stats[:, 0].mean() / nstep > 1        # contacts present most of the time
    assert np.array_equal(out, ref)


@pytest.mark.parametrize("solver", [mb.SOLVER_NEWTON, mb.SOLVER_PGS])
def test_ant_forward_and_rollout_bit_exact(solver):
    """BASELINE config 3 model (models/ant.xml: free root + 4 legs x 2 hinges, condim-3 floor contacts with
    margin, armature, gear-150 motors); native solver Newton"""
    m, b, o = make_pair(ANT, solver, library=hostemu_lib(), nenv=8)
    states = perturbed_states(o, 8, seed=4, height=[0.3, 0.45, 0.6, 0.9], qpos_std=0.2)
    ctrl = np.random.default_rng(6).uniform(-1.5, 1.5, (8, o.size("nu")))
    compare_forward(b, o, states, ctrl, rtol=0, exact=True, check_dual=(solver == mb.SOLVER_PGS))
    nenv, nstep = 8, 150
    s0 = perturbed_states(o, nenv, seed=14, height=[0.35, 0.5, 0.75], qvel_std=0.5, qpos_std=0.15)
    c = np.random.default_rng(15).uniform(-1, 1, (nenv, nstep, o.size("nu")))
    out = b.rollout(s0, c)
    ref, stats, _ = o.rollout(s0, c, nthread=4)
    assert stats[:, 0].mean() / nstep > 0.2      # floor contacts on a good share of the steps
    assert stats[:, 3].sum() == 0
    assert np.array_equal(out, ref)


@pytest.mark.parametrize("solver", [mb.SOLVER_PGS, mb.SOLVER_NEWTON, mb.SOLVER_CG])
def test_frictionloss_rows_bit_exact(solver):
    """dry joint friction (mjCNSTR_FRICTION_DOF rows, nf > 0): Huber cost in the primal solvers, box
    projection in PGS; model = ant with frictionloss on every hinge (cube_3x3x3 relies on these rows)"""
    path = os.path.join(ROOT, "models", "ant_frictionloss.mjb")
    nenv, nstep = 8, 100
    m, b, o = make_pair(path, solver, library=hostemu_lib(), nenv=nenv)
    states = perturbed_states(o, nenv, seed=7, height=[0.3, 0.45, 0.6, 0.9], qpos_std=0.2)
    ctrl1 = np.random.default_rng(8).uniform(-1.5, 1.5, (nenv, o.size("nu")))
    compare_forward(b, o, states, ctrl1, rtol=0, exact=True, check_dual=(solver == mb.SOLVER_PGS))
    assert (b.field("nf")[:, 0] == 14).all()      # 8 hinges + the 6 dofs of the free root (class default)
    s0 = perturbed_states(o, nenv, seed=71, height=[0.35, 0.5, 0.75], qvel_std=0.5, qpos_std=0.15)
    ctrl = np.random.default_rng(72).uniform(-1, 1, (nenv, nstep, o.size("nu")))
    out = b.rollout(s0, ctrl)
    ref, stats, _ = o.rollout(s0, ctrl, nthread=4)
    assert stats[:, 3].sum() == 0
    assert np.array_equal(out, ref)


@pytest.mark.parametrize("solver", [mb.SOLVER_PGS, mb.SOLVER_NEWTON, mb.SOLVER_CG])
def test_constraint_islands_bit_exact(solver):
    """several kinematic trees (models/ant_balls.xml: the ant + two balls + a stick): mj_island partitions
    rows and dofs, and every solver runs once per island (PGS: own shuffle stream, momentum, termination;
    Newton/CG: island-local M, J, Hessian and cost scale), with per-island iteration counts"""
    path = os.path.join(ROOT, "models", "ant_balls.mjb")
    nenv, nstep = 6, 150
    m, b, o = make_pair(path, solver, library=hostemu_lib(), nenv=nenv)
    s0 = perturbed_states(o, nenv, seed=14, height=[0.35, 0.5, 0.75], qvel_std=0.3, qpos_std=0.05)
    ctrl = np.random.default_rng(15).uniform(-1, 1, (nenv, nstep, o.size("nu")))
    out = b.rollout(s0, ctrl)
    ref, stats, _ = o.rollout(s0, ctrl, nthread=4)
    assert stats[:, 3].sum() == 0
    assert np.array_equal(out, ref)
    seen = set()
    for t in (40, 80, 149):     # states with 3-4 islands: every field and the per-island iteration counts
        compare_forward(b, o, out[:, t, :], ctrl[:, t, :], rtol=0, exact=True, check_dual=(solver == mb.SOLVER_PGS))
        seen |= set(b.field("nisland")[:, 0].tolist())
    assert max(seen) >= 3
    # islands disabled: the monolithic solver handles the same forest
    m, b, o = make_pair(path, solver, library=hostemu_lib(), nenv=nenv, disableflags=1 << 18)
    out = b.rollout(s0, ctrl[:, :60])
    ref, _, _ = o.rollout(s0, ctrl[:, :60], nthread=4)
    assert np.array_equal(out, ref)


@pytest.mark.parametrize("model", [HUMANOID, ANT])
def test_cg_solver_bit_exact(model):
    """mj_solCG: the primal machinery with M-preconditioning and the Hager-Zhang direction update"""
    nenv, nstep = 8, 80
    m, b, o = make_pair(model, mb.SOLVER_CG, library=hostemu_lib(), nenv=nenv)
    states = perturbed_states(o, nenv, seed=3, height=[0.25, 0.4, 0.9, 1.3])
    ctrl1 = np.random.default_rng(5).uniform(-1.5, 1.5, (nenv, o.size("nu")))
    compare_forward(b, o, states, ctrl1, rtol=0, exact=True, check_dual=False)
    s0 = perturbed_states(o, nenv, seed=61, height=[0.3, 0.5, 0.8], qvel_std=0.5, qpos_std=0.15)
    ctrl = np.random.default_rng(62).uniform(-1, 1, (nenv, nstep, o.size("nu")))
    out = b.rollout(s0, ctrl)
    ref, stats, _ = o.rollout(s0, ctrl, nthread=4)
    assert stats[:, 0].sum() > 0 and stats[:, 3].sum() == 0
    assert np.array_equal(out, ref)


@pytest.mark.parametrize("model,solver", [(HUMANOID, mb.SOLVER_PGS), (HUMANOID, mb.SOLVER_NEWTON), (ANT, mb.SOLVER_NEWTON)])
def test_rk4_rollout_bit_exact(model, solver):
    """integrator = RK4 (mj_RungeKutta): 4 forward passes per step; BASELINE config 5's integrator"""
    nenv, nstep = 6, 60
    m, b, o = make_pair(model, solver, library=hostemu_lib(), nenv=nenv, integrator=mb.INT_RK4)
    s0 = perturbed_states(o, nenv, seed=51, height=[0.3, 0.5, 0.8], qvel_std=0.5, qpos_std=0.15)
    ctrl = np.random.default_rng(52).uniform(-1, 1, (nenv, nstep, o.size("nu")))
    out = b.rollout(s0, ctrl)
    ref, stats, _ = o.rollout(s0, ctrl, nthread=3)
    assert stats[:, 0].sum() > 0 and stats[:, 3].sum() == 0
    assert np.array_equal(out, ref)
    # a bad state inside a rollout: reset + finish the step (RK4 is a multi-launch step), then pad
    s0[2, 5] = np.inf
    out = b.rollout(s0, ctrl[:, :8])
    ref, stats, _ = o.rollout(s0, ctrl[:, :8], nthread=1)
    assert stats[2, 3] > 0
    assert np.array_equal(out, ref)


def test_option_variants_bit_exact():
    """disable flags and solver options that change control flow on the path"""
    DSBL_WARMSTART, DSBL_CLAMPCTRL, DSBL_EULERDAMP, DSBL_REFSAFE, DSBL_FILTERPARENT = 1 << 9, 1 << 8, 1 << 15, 1 << 12, 1 << 10
    DSBL_CONTACT, DSBL_LIMIT, DSBL_GRAVITY, DSBL_SPRING, DSBL_DAMPER, DSBL_CONSTRAINT = 1 << 4, 1 << 3, 1 << 7, 1 << 5, 1 << 6, 1
    nenv, nstep = 4, 40
    for flags, extra in [(DSBL_WARMSTART, {}), (DSBL_CLAMPCTRL | DSBL_EULERDAMP, {}), (DSBL_REFSAFE | DSBL_FILTERPARENT, {}),
                         (DSBL_CONTACT, {}), (DSBL_LIMIT | DSBL_GRAVITY, {}), (DSBL_SPRING, {}), (DSBL_DAMPER, {}),
                         (DSBL_CONSTRAINT, {}), (0, {"iterations": 7}), (0, {"tolerance": 1e-4, "impratio": 3.0}),
                         (0, {"timestep": 0.002})]:
        m, b, o = make_pair(HUMANOID, mb.SOLVER_PGS, library=hostemu_lib(), nenv=nenv, disableflags=flags, **extra)
        s0 = perturbed_states(o, nenv, seed=31, height=[0.25, 0.6], qvel_std=0.3, qpos_std=0.15)
        ctrl = np.random.default_rng(32).uniform(-1.2, 1.2, (nenv, nstep, o.size("nu")))
        out = b.rollout(s0, ctrl)
        ref, _, _ = o.rollout(s0, ctrl, nthread=2)
        assert np.array_equal(out, ref), (flags, extra, np.abs(out - ref).max())


def test_bad_state_warning_and_padding():
    """rollout.cc:127-155: an env that raises a warning stops stepping and pads its outputs;
    mj_checkPos auto-resets to qpos0 (engine_forward.c:54-69)"""
    nenv, nstep = 3, 6
    m, b, o = make_pair(HUMANOID, mb.SOLVER_PGS, library=hostemu_lib(), nenv=nenv)
    o.reset()
    s0 = np.tile(o.get_state(), (nenv, 1))
    s0[1, 3] = np.nan                              # bad qpos in env 1
    ctrl = np.zeros((nenv, nstep, o.size("nu")))
    out = b.rollout(s0, ctrl)
    ref, stats, _ = o.rollout(s0, ctrl, nthread=1)
    assert np.array_equal(out[0], ref[0]) and np.array_equal(out[2], ref[2])
    assert np.array_equal(out[1], ref[1])          # reset + one step, then padded
    w = b.warnings()
    assert w[1, 3] == 1 and w[0].sum() == 0


def test_state_io_and_argument_errors():
    m = mb.Model(HUMANOID, library=hostemu_lib())
    b = mb.Batch(m, 5)
    assert b.state_size() == 1 + 28 + 27
    s = np.random.default_rng(0).normal(size=(5, b.state_size()))
    b.set_state(s)
    assert np.array_equal(b.get_state(), s)
    with pytest.raises(ValueError):
        b.set_state(s[:, :-1])
    with pytest.raises(ValueError):
        b.rollout(s, np.zeros((5, 3, 7)))
    assert b.state_size(1 << 8) == 6 * m.size("nbody")   # xfrc_applied
    with pytest.raises(mb.MjbError):
        b.state_size(1 << 20)       # not an mjtState bit


def test_unsupported_models_are_refused():
    m = mb.Model(HUMANOID, library=hostemu_lib())
    m.set_option("cone", 1)
    with pytest.raises(mb.MjbError, match="elliptic"):
        mb.Batch(m, 1)
    m.set_option("cone", 0)
    m.set_option("integrator", 2)
    with pytest.raises(mb.MjbError, match="implicit"):
        mb.Batch(m, 1)


@pytest.mark.parametrize("model", ["humanoid", "ant_act", "ant_sensors", "humanoid+energy"])
def test_mjdata_bridge_matches_mj_step(model):
    """mjb_step_mjdata: the reference's per-mjData loop `for k: mj_step(m, d[k])` as one call.  Two sets
    of the reference's own mjData objects start identical; one is stepped by the reference engine, the other
    through the bridge; every fixed-size member the path computes must then be identical"""
    from oracle_util import Oracle
    nenv, nstep = 3, 25
    energy_flag = model.endswith("+energy")       # mjENBL_ENERGY: mj_energyPos / mj_energyVel every step
    path = os.path.join(ROOT, "models", model.split("+")[0] + ".mjb")
    ref = [Oracle(path) for _ in range(nenv)]
    ours = [Oracle(path) for _ in range(nenv)]
    m = mb.Model(path, library=hostemu_lib())
    m.set_option("solver", mb.SOLVER_NEWTON)
    if energy_flag:
        m.set_option("enableflags", 2)
        for o in ref + ours:
            o.set_opt("enableflags", 2)
    b = mb.Batch(m, nenv, nconmax=64, njmax=200)
    states = perturbed_states(ref[0], nenv, seed=33, height=[0.3, 0.5, 0.9], qvel_std=0.4, qpos_std=0.1)
    rng = np.random.default_rng(34)
    for e in range(nenv):
        for o in (ref[e], ours[e]):
            o.set_opt("solver", mb.SOLVER_NEWTON)
            o.reset()
            o.set_state(states[e])
    fields = ["qpos", "qvel", "qacc", "qacc_warmstart", "xpos", "xquat", "xmat", "xipos", "geom_xpos", "geom_xmat",
              "subtree_com", "cinert", "cdof", "crb", "M", "qLD", "qLDiagInv", "cvel", "cdof_dot", "qfrc_bias",
              "qfrc_passive", "actuator_force", "qfrc_actuator", "qfrc_smooth", "qacc_smooth", "qfrc_constraint", "energy"]
    if ref[0].size("na"):      # stateful actuators: act goes in and comes back, act_dot comes back
        fields += ["act", "act_dot", "actuator_length", "actuator_velocity"]
    if ref[0].size("nsensordata"):   # sensors, sites and the rnePostConstraint / subtreeVel outputs the sensors need
        fields += ["sensordata", "site_xpos", "site_xmat", "cacc", "cfrc_int", "cfrc_ext", "subtree_linvel", "subtree_angmom"]
    for t in range(nstep):
        for e in range(nenv):
            c = rng.uniform(-1, 1, ref[e].size("nu"))
            ref[e].dfield("ctrl")[:] = c
            ours[e].dfield("ctrl")[:] = c
            ref[e].step()
        b.step_mjdata([o.d for o in ours])
        for e in range(nenv):
            assert ref[e].scalar("time") == ours[e].scalar("time")
            # arena members are not materialised in the bridged mjData: its counts read 0, the batch has the real ones
            assert ours[e].scalar("ncon") == 0 and ours[e].scalar("nefc") == 0
            assert ref[e].scalar("ncon") == b.field("ncon")[e, 0] and ref[e].scalar("nefc") == b.field("nefc")[e, 0]
            for f in fields:
                assert np.array_equal(np.array(ref[e].dfield(f)), np.array(ours[e].dfield(f))), (t, e, f)


@pytest.mark.parametrize("solver,integrator", [(mb.SOLVER_NEWTON, mb.INT_EULER), (mb.SOLVER_PGS, mb.INT_EULER),
                                               (mb.SOLVER_NEWTON, mb.INT_RK4)])
def test_sensors_bit_exact(solver, integrator):
    """sensordata (engine_sensor.c): one sensor of every supported type — joint / tendon / actuator
    position, velocity and force, limit distance / velocity / force, ball-joint quaternion and rate, frame
    position / axes / quaternion / linear and angular velocity (absolute and relative to a reference
    frame, objects xbody / body / geom), subtree com, clock, with and without cutoff — recorded by
    rollout() like python/mujoco/rollout.cc:160-170; RK4 evaluates sensors in the first forward only"""
    from oracle_util import Oracle
    path = os.path.join(ROOT, "models", "ant_sensors.mjb")
    nenv, nstep = 4, 60
    m = mb.Model(path, library=hostemu_lib())
    m.set_option("solver", solver)
    m.set_option("integrator", integrator)
    b = mb.Batch(m, nenv, nconmax=48, njmax=128)
    oracles = [Oracle(path) for _ in range(nenv)]
    rough = integrator == mb.INT_EULER     # explicit RK4 at dt = 0.01 diverges from the rougher states
    s0 = perturbed_states(oracles[0], nenv, seed=91, height=[0.35, 0.5, 0.75], qvel_std=0.5 if rough else 0.1,
                          qpos_std=0.15 if rough else 0.03)
    ctrl = np.random.default_rng(92).uniform(-1, 1, (nenv, nstep, oracles[0].size("nu")))
    out, sens = b.rollout(s0, ctrl, return_sensordata=True)
    ns = oracles[0].size("nsensordata")
    assert sens.shape == (nenv, nstep, ns) and ns == 271
    for e, o in enumerate(oracles):
        o.set_opt("solver", solver)
        o.set_opt("integrator", integrator)
        o.reset()
        o.set_state(s0[e])
        for t in range(nstep):
            o.dfield("ctrl")[:] = ctrl[e, t]
            o.step()
            ref = np.array(o.dfield("sensordata"))
            assert np.array_equal(sens[e, t], ref), (e, t, np.argwhere(sens[e, t] != ref).ravel()[:8])
            assert np.array_equal(out[e, t], o.get_state())
    assert np.abs(sens).max() > 0 and (b.warnings() == 0).all()


@pytest.mark.parametrize("model", ["humanoid", "ant", "ant_servo"])
@pytest.mark.parametrize("solver", [mb.SOLVER_NEWTON, mb.SOLVER_PGS])
def test_implicitfast_bit_exact(model, solver):
    """integrator = implicitfast (mj_implicitSkip): qH = M - h d(qfrc_actuator + qfrc_passive)/d(qvel).
    ant_servo has velocity-dependent actuators (position / velocity servos, an affine general actuator,
    two of them force-limited so the clamp test matters) and damped tendons, one along a kinematic chain
    and one across unrelated legs (whose cross terms fall outside the tree sparsity and are dropped)"""
    path = os.path.join(ROOT, "models", model + ".mjb")
    nenv, nstep = 6, 150
    m, b, o = make_pair(path, solver, library=hostemu_lib(), nenv=nenv, integrator=mb.INT_IMPLICITFAST)
    s0 = perturbed_states(o, nenv, seed=14, height=[0.35, 0.5, 0.75], qvel_std=0.3, qpos_std=0.05)
    ctrl = np.random.default_rng(15).uniform(-1, 1, (nenv, nstep, o.size("nu")))
    out = b.rollout(s0, ctrl)
    ref, stats, _ = o.rollout(s0, ctrl, nthread=4)
    assert stats[:, 3].sum() == 0
    assert np.array_equal(out, ref)


@pytest.mark.parametrize("model,solver", [("ant_balls", mb.SOLVER_NEWTON), ("ant_balls", mb.SOLVER_PGS), ("boxes", mb.SOLVER_NEWTON)])
def test_implicitfast_standalone_free_bodies_bit_exact(model, solver):
    """implicitfast with bodies that are a whole tree by themselves (one free joint, no children): their six
    accelerations come from the local unsymmetric solve of engine_forward.c:1742-1762 - M - h dqfrc_smooth/dqvel with
    the gyroscopic (bias) block of mjd_freeBias_vel, partially pivoted 6x6 LU - instead of the symmetric qH solve.
    Spinning balls next to the ant (islands) and the tumbling boxes / cylinders of models/boxes.xml."""
    path = os.path.join(ROOT, "models", model + ".mjb")
    nenv, nstep = 6, 80
    m, b, o = make_pair(path, solver, library=hostemu_lib(), nenv=nenv, nconmax=96, njmax=400, integrator=mb.INT_IMPLICITFAST)
    nq, nv = o.size("nq"), o.size("nv")
    rng = np.random.default_rng(41)
    if model == "boxes":
        o.reset()
        s0 = np.tile(o.get_state(), (nenv, 1))
        for e in range(nenv):
            for k in range(9):
                s0[e, 1 + 7 * k + 2] += rng.uniform(-0.05, 0.3)
                s0[e, 1 + 7 * k + 3:1 + 7 * k + 7] = rng.normal(size=4)
            s0[e, 1 + nq:] = rng.normal(0, 2.5, nv)        # fast spins: the gyroscopic block matters
    else:
        s0 = perturbed_states(o, nenv, seed=42, height=[0.35, 0.5, 0.75], qvel_std=2.0, qpos_std=0.1)
    ctrl = rng.uniform(-1, 1, (nenv, nstep, o.size("nu")))
    out = b.rollout(s0, ctrl)
    ref, stats, _ = o.rollout(s0, ctrl, nthread=4)
    assert stats[:, 3].sum() == 0
    assert np.array_equal(out, ref)
    # the local solve changes the result: Euler-style symmetric-only integration of the same model differs
    m2, b2, o2 = make_pair(path, solver, library=hostemu_lib(), nenv=nenv, nconmax=96, njmax=400, integrator=mb.INT_EULER)
    assert not np.array_equal(b2.rollout(s0, ctrl), out)


@pytest.mark.parametrize("solver", [mb.SOLVER_PGS, mb.SOLVER_NEWTON, mb.SOLVER_CG])
def test_equality_couplings_bit_exact(solver):
    """joint / tendon equality constraints (mj_instantiateEquality, scalar couplings with the quartic
    polynomial): equality, frictionloss, limit and contact rows in one problem (models/ant_equality.xml:
    5 equalities, one of them inactive)"""
    path = os.path.join(ROOT, "models", "ant_equality.mjb")
    nenv, nstep = 6, 150
    m, b, o = make_pair(path, solver, library=hostemu_lib(), nenv=nenv)
    states = perturbed_states(o, nenv, seed=4, height=[0.3, 0.45, 0.6], qpos_std=0.1)
    ctrl1 = np.random.default_rng(6).uniform(-1, 1, (nenv, o.size("nu")))
    compare_forward(b, o, states, ctrl1, rtol=0, exact=True, check_dual=(solver == mb.SOLVER_PGS))
    assert (b.field("ne")[:, 0] == 4).all()
    s0 = perturbed_states(o, nenv, seed=14, height=[0.35, 0.5, 0.75], qvel_std=0.3, qpos_std=0.05)
    ctrl = np.random.default_rng(15).uniform(-1, 1, (nenv, nstep, o.size("nu")))
    out = b.rollout(s0, ctrl)
    ref, stats, _ = o.rollout(s0, ctrl, nthread=4)
    assert stats[:, 3].sum() == 0
    assert np.array_equal(out, ref)
    # mjDSBL_EQUALITY drops the rows
    m, b, o = make_pair(path, solver, library=hostemu_lib(), nenv=nenv, disableflags=1 << 1)
    out = b.rollout(s0, ctrl[:, :40])
    ref, _, _ = o.rollout(s0, ctrl[:, :40], nthread=4)
    assert np.array_equal(out, ref) and (b.field("ne")[:, 0] == 0).all()


@pytest.mark.parametrize("solver", [mb.SOLVER_PGS, mb.SOLVER_NEWTON, mb.SOLVER_CG])
def test_connect_equality_bit_exact(solver):
    """connect (ball-joint) equalities: 3 rows with J = jac(anchor 0) - jac(anchor 1), common impedance from
    |pos|, the Jdot*v correction of aref (mj_Jdotv / mj_jacDot), an equality row joining two kinematic
    trees into one island, a closed loop inside a tree, and the connect forces entering
    mj_rnePostConstraint (accelerometer / force / frame acceleration sensors) - models/ant_connect.xml"""
    from oracle_util import Oracle
    path = os.path.join(ROOT, "models", "ant_connect.mjb")
    nenv, nstep = 4, 100
    m, b, o = make_pair(path, solver, library=hostemu_lib(), nenv=nenv)
    states = perturbed_states(o, nenv, seed=4, height=[0.3, 0.45, 0.6], qpos_std=0.1)
    ctrl1 = np.random.default_rng(6).uniform(-1, 1, (nenv, o.size("nu")))
    compare_forward(b, o, states, ctrl1, rtol=0, exact=True, check_dual=(solver == mb.SOLVER_PGS))
    assert (b.field("ne")[:, 0] == 6).all() and (b.field("nisland")[:, 0] == 1).all()
    s0 = perturbed_states(o, nenv, seed=14, height=[0.35, 0.5, 0.75], qvel_std=0.3, qpos_std=0.05)
    ctrl = np.random.default_rng(15).uniform(-1, 1, (nenv, nstep, o.size("nu")))
    out, sens = b.rollout(s0, ctrl, return_sensordata=True)
    for e in range(nenv):
        oe = Oracle(path)
        oe.set_opt("solver", solver)
        oe.reset()
        oe.set_state(s0[e])
        for t in range(nstep):
            oe.dfield("ctrl")[:] = ctrl[e, t]
            oe.step()
            assert np.array_equal(out[e, t], oe.get_state()), (e, t)
            assert np.array_equal(sens[e, t], np.array(oe.dfield("sensordata"))), (e, t)


@pytest.mark.parametrize("solver", [mb.SOLVER_PGS, mb.SOLVER_NEWTON, mb.SOLVER_CG])
def test_weld_equality_bit_exact(solver):
    """weld equalities: 3 translational rows (connect at the weld anchors) + 3 rotational rows
    0.5*neg(q1)*(jacr0-jacr1)*q0*relpose*torquescale, impedance from the 6-row |pos|, translational and
    rotational Jdot*v corrections, a weld to the world body, weld forces and torques in
    mj_rnePostConstraint (force / torque sensors) - models/ant_weld.xml"""
    from oracle_util import Oracle
    path = os.path.join(ROOT, "models", "ant_weld.mjb")
    nenv, nstep = 4, 100
    m, b, o = make_pair(path, solver, library=hostemu_lib(), nenv=nenv)
    states = perturbed_states(o, nenv, seed=4, height=[0.3, 0.45, 0.6], qpos_std=0.1)
    ctrl1 = np.random.default_rng(6).uniform(-1, 1, (nenv, o.size("nu")))
    compare_forward(b, o, states, ctrl1, rtol=0, exact=True, check_dual=(solver == mb.SOLVER_PGS))
    assert (b.field("ne")[:, 0] == 15).all() and (b.field("nisland")[:, 0] == 2).all()
    s0 = perturbed_states(o, nenv, seed=14, height=[0.35, 0.5, 0.75], qvel_std=0.3, qpos_std=0.05)
    ctrl = np.random.default_rng(15).uniform(-1, 1, (nenv, nstep, o.size("nu")))
    out, sens = b.rollout(s0, ctrl, return_sensordata=True)
    for e in range(nenv):
        oe = Oracle(path)
        oe.set_opt("solver", solver)
        oe.reset()
        oe.set_state(s0[e])
        for t in range(nstep):
            oe.dfield("ctrl")[:] = ctrl[e, t]
            oe.step()
            assert np.array_equal(out[e, t], oe.get_state()), (e, t)
            assert np.array_equal(sens[e, t], np.array(oe.dfield("sensordata"))), (e, t)


@pytest.mark.parametrize("model,integrator", [("ant_act", mb.INT_EULER), ("ant_act", mb.INT_RK4),
                                              ("ant_act_nomuscle", mb.INT_EULER), ("ant_act_nomuscle", mb.INT_RK4),
                                              ("ant_act_nomuscle", mb.INT_IMPLICITFAST)])
@pytest.mark.parametrize("solver", [mb.SOLVER_NEWTON, mb.SOLVER_PGS])
def test_stateful_actuators_bit_exact(model, integrator, solver):
    """activation state `act` (part of mjSTATE_FULLPHYSICS): filter, filterexact (exact exponential update),
    integrator with actrange clamp, actearly, muscles (FLV gain, passive bias, activation dynamics), and
    tendon transmissions (moment row = ten_J row * gear) - models/ant_act*.xml; the state vector returned by
    rollout carries act, RK4 integrates it with act_dot, implicitfast differentiates gain*act"""
    path = os.path.join(ROOT, "models", model + ".mjb")
    nenv, nstep = 6, 120
    m, b, o = make_pair(path, solver, library=hostemu_lib(), nenv=nenv, integrator=integrator)
    na, nq, nv = o.size("na"), o.size("nq"), o.size("nv")
    assert na > 0 and b.state_size() == 1 + nq + nv + na
    s0 = perturbed_states(o, nenv, seed=14, height=[0.35, 0.5, 0.75], qvel_std=0.3, qpos_std=0.05)
    s0[:, 1 + nq + nv:] = np.random.default_rng(3).uniform(-0.3, 0.8, (nenv, na))
    ctrl = np.random.default_rng(15).uniform(-1, 1, (nenv, nstep, o.size("nu")))
    compare_forward(b, o, s0, ctrl[:, 0], rtol=0, exact=True, check_dual=(solver == mb.SOLVER_PGS))
    out = b.rollout(s0, ctrl)
    ref, stats, _ = o.rollout(s0, ctrl, nthread=4)
    assert stats[:, 3].sum() == 0
    assert np.array_equal(out, ref)
    assert np.abs(out[:, -1, 1 + nq + nv:]).max() > 0


@pytest.mark.parametrize("solver", [mb.SOLVER_NEWTON, mb.SOLVER_PGS])
def test_mocap_body_bit_exact(solver):
    """mocap body (pose read from mjData.mocap_pos / mocap_quat in mj_kinematics, unnormalised quaternions
    included) dragging a free body through a soft weld - models/ant_mocap.xml.  The mocap trajectory enters
    rollout() through control_spec = CTRL | MOCAP_POS | MOCAP_QUAT like in the reference; without those bits
    the model's pose is used"""
    from oracle_util import Oracle
    path = os.path.join(ROOT, "models", "ant_mocap.mjb")
    nenv, nstep = 3, 80
    m, b, o = make_pair(path, solver, library=hostemu_lib(), nenv=nenv)
    nu = o.size("nu")
    s0 = perturbed_states(o, nenv, seed=14, height=[0.35, 0.5, 0.75], qvel_std=0.3, qpos_std=0.05)
    rng = np.random.default_rng(5)
    ctrl = rng.uniform(-1, 1, (nenv, nstep, nu))
    t = np.arange(nstep)[None, :, None] * 0.01
    mpos = np.array([1.5, 0.2, 0.7]) + 0.3 * np.sin(3 * t + rng.uniform(0, 6, (nenv, 1, 3)))
    mquat = np.array([0.98, 0.1, 0.1, 0.12]) + 0.2 * np.cos(2 * t + rng.uniform(0, 6, (nenv, 1, 4)))
    spec = mb.STATE_CTRL | mb.STATE_MOCAP_POS | mb.STATE_MOCAP_QUAT
    assert b.state_size(spec) == nu + 7
    out, sens = b.rollout(s0, np.concatenate([ctrl, mpos, mquat], axis=2), control_spec=spec, return_sensordata=True)
    out_default, sens_default = b.rollout(s0, ctrl, return_sensordata=True)          # mocap pose falls back to the model's
    for e in range(nenv):
        for moving, got, gsens in ((True, out, sens), (False, out_default, sens_default)):
            oe = Oracle(path)
            oe.set_opt("solver", solver)
            oe.reset()
            oe.set_state(s0[e])
            for k in range(nstep):
                oe.dfield("ctrl")[:] = ctrl[e, k]
                if moving:
                    oe.dfield("mocap_pos")[:] = mpos[e, k]
                    oe.dfield("mocap_quat")[:] = mquat[e, k]
                oe.step()
                assert np.array_equal(got[e, k], oe.get_state()), (e, k, moving)
                # incl. subtreelinvel / subtreeangmom (mj_subtreeVel) of moving, static and leaf subtrees
                assert np.array_equal(gsens[e, k], np.array(oe.dfield("sensordata"))), (e, k, moving)
    assert not np.array_equal(out, out_default)


@pytest.mark.parametrize("solver,integrator", [(mb.SOLVER_NEWTON, mb.INT_EULER), (mb.SOLVER_PGS, mb.INT_EULER),
                                               (mb.SOLVER_NEWTON, mb.INT_RK4)])
def test_fluid_forces_bit_exact(solver, integrator):
    """medium with density, viscosity and wind (inertia-box fluid model, mj_inertiaBoxFluidModel): per-body
    viscous + quadratic drag wrench from the body-local velocity, mapped by mj_applyFT - models/ant_fluid.xml"""
    path = os.path.join(ROOT, "models", "ant_fluid.mjb")
    nenv, nstep = 6, 120
    m, b, o = make_pair(path, solver, library=hostemu_lib(), nenv=nenv, integrator=integrator)
    s0 = perturbed_states(o, nenv, seed=14, height=[0.35, 0.5, 0.75], qvel_std=1.5, qpos_std=0.05)
    ctrl = np.random.default_rng(15).uniform(-1, 1, (nenv, nstep, o.size("nu")))
    compare_forward(b, o, s0, ctrl[:, 0], rtol=0, exact=True, check_dual=(solver == mb.SOLVER_PGS))
    out = b.rollout(s0, ctrl)
    ref, stats, _ = o.rollout(s0, ctrl, nthread=4)
    assert stats[:, 3].sum() == 0
    assert np.array_equal(out, ref)
    assert np.abs(b.field("qfrc_fluid")).max() > 0


@pytest.mark.parametrize("model,solver", [("ant_weld", mb.SOLVER_NEWTON), ("humanoid", mb.SOLVER_PGS)])
def test_xfrc_applied_bit_exact(model, solver):
    """Cartesian perturbation wrenches mjData.xfrc_applied (mj_xfrcAccumulate -> qfrc_smooth; cfrc_ext in
    mj_rnePostConstraint, seen by the force / torque / accelerometer sensors of ant_weld) as a rollout input
    (control_spec = CTRL | XFRC_APPLIED) and through the mjData bridge"""
    from oracle_util import Oracle
    path = os.path.join(ROOT, "models", model + ".mjb")
    nenv, nstep = 3, 60
    m, b, o = make_pair(path, solver, library=hostemu_lib(), nenv=nenv)
    nu, nbody, nsens = o.size("nu"), o.size("nbody"), o.size("nsensordata")
    s0 = perturbed_states(o, nenv, seed=14, height=[0.5, 0.9, 1.3] if model == "humanoid" else [0.35, 0.5, 0.75], qvel_std=0.3, qpos_std=0.05)
    rng = np.random.default_rng(5)
    ctrl = rng.uniform(-1, 1, (nenv, nstep, nu))
    xfrc = np.zeros((nenv, nstep, nbody, 6))
    xfrc[:, :, 1, :] = rng.normal(0, 4, (nenv, nstep, 6))            # torso: force and torque every step
    xfrc[:, 20:40, nbody - 1, :3] = rng.normal(0, 2, (nenv, 20, 3))  # last body: a force pulse
    xfrc[:, :, 3, 2] = 1.5                                            # constant lift on body 3 (zero components skipped)
    spec = mb.STATE_CTRL | mb.STATE_XFRC_APPLIED
    control = np.concatenate([ctrl, xfrc.reshape(nenv, nstep, -1)], axis=2)
    if nsens:
        out, sens = b.rollout(s0, control, control_spec=spec, return_sensordata=True)
    else:
        out = b.rollout(s0, control, control_spec=spec)
    for e in range(nenv):
        oe = Oracle(path)
        oe.set_opt("solver", solver)
        oe.reset()
        oe.set_state(s0[e])
        for k in range(nstep):
            oe.dfield("ctrl")[:] = ctrl[e, k]
            oe.dfield("xfrc_applied")[:] = xfrc[e, k]
            oe.step()
            assert np.array_equal(out[e, k], oe.get_state()), (e, k)
            if nsens:
                assert np.array_equal(sens[e, k], np.array(oe.dfield("sensordata"))), (e, k)
    # without the bit the perturbations are cleared again (rollout.cc:92-94)
    out0 = b.rollout(s0, ctrl)
    ref0, _, _ = o.rollout(s0, ctrl, nthread=2)
    assert np.array_equal(out0, ref0) and not np.array_equal(out0, out)
    # bridge: xfrc_applied is read from the caller's mjData
    ours = [Oracle(path) for _ in range(nenv)]
    refs = [Oracle(path) for _ in range(nenv)]
    b2 = mb.Batch(m, nenv, nconmax=48, njmax=160)
    for e in range(nenv):
        for x in (ours[e], refs[e]):
            x.set_opt("solver", solver); x.reset(); x.set_state(s0[e])
    for k in range(10):
        for e in range(nenv):
            for x in (ours[e], refs[e]):
                x.dfield("ctrl")[:] = ctrl[e, k]
                x.dfield("xfrc_applied")[:] = xfrc[e, k]
            refs[e].step()
        b2.step_mjdata([x.d for x in ours])
        for e in range(nenv):
            assert np.array_equal(np.array(ours[e].dfield("qpos")), np.array(refs[e].dfield("qpos"))), (e, k)
            assert np.array_equal(np.array(ours[e].dfield("qacc")), np.array(refs[e].dfield("qacc"))), (e, k)


def test_eq_active_runtime_switch_bit_exact():
    """mjData.eq_active as a runtime input: equalities switched off and on during a rollout through
    control_spec = CTRL | EQ_ACTIVE (values 0 / 1 as doubles, like mj_setState), default = eq_active0"""
    from oracle_util import Oracle
    path = os.path.join(ROOT, "models", "ant_weld.mjb")
    nenv, nstep = 3, 60
    m, b, o = make_pair(path, mb.SOLVER_NEWTON, library=hostemu_lib(), nenv=nenv)
    nu, neq = o.size("nu"), o.size("neq")
    s0 = perturbed_states(o, nenv, seed=14, height=[0.35, 0.5, 0.75], qvel_std=0.3, qpos_std=0.05)
    rng = np.random.default_rng(5)
    ctrl = rng.uniform(-1, 1, (nenv, nstep, nu))
    act = (rng.uniform(0, 1, (nenv, nstep // 10, neq)) > 0.4).astype(np.float64).repeat(10, axis=1)   # switches every 10 steps
    spec = mb.STATE_CTRL | mb.STATE_EQ_ACTIVE
    assert b.state_size(spec) == nu + neq
    out = b.rollout(s0, np.concatenate([ctrl, act], axis=2), control_spec=spec)
    for e in range(nenv):
        oe = Oracle(path)
        oe.set_opt("solver", mb.SOLVER_NEWTON)
        oe.reset()
        oe.set_state(s0[e])
        for k in range(nstep):
            oe.dfield("ctrl")[:] = ctrl[e, k]
            oe.dfield("eq_active")[:] = act[e, k].astype(np.uint8)
            oe.step()
            assert np.array_equal(out[e, k], oe.get_state()), (e, k)
    ref0, _, _ = o.rollout(s0, ctrl, nthread=2)          # default: every equality as in the model
    assert np.array_equal(b.rollout(s0, ctrl), ref0) and not np.array_equal(ref0, out)


@pytest.mark.parametrize("solver", [mb.SOLVER_PGS, mb.SOLVER_NEWTON, mb.SOLVER_CG])
def test_condim_4_and_6_bit_exact(solver):
    """torsional and rolling friction (condim 4 / 6): pyramids with 6 / 10 edges whose extra rows use the
    rotational contact Jacobian (mj_contactJacobian), diagApprox with the rotational weights, contact torques in
    the decoded contact force - models/ant_condim.xml"""
    path = os.path.join(ROOT, "models", "ant_condim.mjb")
    nenv, nstep = 6, 120
    m, b, o = make_pair(path, solver, library=hostemu_lib(), nenv=nenv, nconmax=48, njmax=220)
    s0 = perturbed_states(o, nenv, seed=14, height=[0.35, 0.5, 0.75], qvel_std=0.8, qpos_std=0.05)
    ctrl = np.random.default_rng(15).uniform(-1, 1, (nenv, nstep, o.size("nu")))
    compare_forward(b, o, s0, ctrl[:, 0], rtol=0, exact=True, check_dual=(solver == mb.SOLVER_PGS))
    out, sens = b.rollout(s0, ctrl, return_sensordata=True)
    ref, stats, _ = o.rollout(s0, ctrl, nthread=4)
    assert stats[:, 3].sum() == 0
    assert np.array_equal(out, ref)
    assert b.field("con_dim").max() >= 4
    from oracle_util import Oracle      # contact torques of the 4 / 6-dim contacts reach cfrc_ext (force / torque sensors)
    oe = Oracle(path)
    oe.set_opt("solver", solver)
    oe.reset()
    oe.set_state(s0[0])
    for t in range(nstep):
        oe.dfield("ctrl")[:] = ctrl[0, t]
        oe.step()
        assert np.array_equal(sens[0, t], np.array(oe.dfield("sensordata"))), t


@pytest.mark.parametrize("solver", [mb.SOLVER_NEWTON, mb.SOLVER_PGS])
def test_touch_zones_bit_exact(solver):
    """touch sensors with every zone shape (sphere, box, capsule, ellipsoid, cylinder: mju_rayGeom against the
    site volume) - models/ant_touch.xml (the mocap model plus the extra zones)"""
    from oracle_util import Oracle
    path = os.path.join(ROOT, "models", "ant_touch.mjb")
    nenv, nstep = 3, 80
    m, b, o = make_pair(path, solver, library=hostemu_lib(), nenv=nenv)
    s0 = perturbed_states(o, nenv, seed=14, height=[0.35, 0.5, 0.75], qvel_std=0.3, qpos_std=0.05)
    ctrl = np.random.default_rng(5).uniform(-1, 1, (nenv, nstep, o.size("nu")))
    out, sens = b.rollout(s0, ctrl, return_sensordata=True)
    hits = np.zeros(6, dtype=int)
    for e in range(nenv):
        oe = Oracle(path)
        oe.set_opt("solver", solver)
        oe.reset()
        oe.set_state(s0[e])
        for k in range(nstep):
            oe.dfield("ctrl")[:] = ctrl[e, k]
            oe.step()
            r = np.array(oe.dfield("sensordata"))
            assert np.array_equal(out[e, k], oe.get_state()) and np.array_equal(sens[e, k], r), (e, k)
            hits += (r[12:18] != 0)
    assert (hits > 0).all(), hits      # every zone shape saw a contact


def test_step_host_writes_controls_after_a_warning():
    """a benign warning (contact buffer full) must not freeze the controls of the per-step API: only the rollout
    path applies the stop-on-warning rule (rollout.cc:127-155)"""
    m = mb.Model(ANT, library=hostemu_lib())
    b = mb.Batch(m, 2, nconmax=1, njmax=64)
    b.reset()
    s = b.get_state()
    s[:, 3] = 0.1                                   # low: several contacts, more than nconmax = 1
    b.set_state(s)
    nu, ns = m.size("nu"), b.state_size()
    out = np.zeros((2, ns))
    b.step_host(np.full((2, nu), 0.5), out)
    assert b.warnings()[:, 1].min() >= 1            # mjWARN_CONTACTFULL raised
    b.step_host(np.full((2, nu), 0.6), out)
    assert np.array_equal(b.field("ctrl"), np.full((2, nu), 0.6))


def test_rollout_rejects_non_input_control_bits():
    m = mb.Model(ANT, library=hostemu_lib())
    b = mb.Batch(m, 2)
    s0 = b.get_state()
    nu, nv = m.size("nu"), m.size("nv")
    with pytest.raises(mb.MjbError, match="control_spec"):
        b.rollout(s0, np.zeros((2, 3, nu + nv)), control_spec=mb.STATE_CTRL | mb.STATE_WARMSTART)


def test_single_environment_symbols_match_the_reference():
    """mj_step / mj_forward / mj_forwardSkip / mj_step1 + mj_step2 exported under the reference's names
    (include/mujoco/mujoco.h:189-204): the loop of sample/testspeed.cc:123 run through them on the reference's own
    mjData gives the reference's trajectory (host build of the same source; the GPU test repeats it on the device)"""
    import ctypes as C
    from oracle_util import Oracle
    lib = C.CDLL(HOSTEMU)
    for name in ("mj_step", "mj_forward", "mj_step1", "mj_step2"):
        getattr(lib, name).argtypes = [C.c_void_p, C.c_void_p]
        getattr(lib, name).restype = None
    lib.mj_forwardSkip.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int]
    lib.mjb_forget_model.argtypes = [C.c_void_p]
    ref, ours = Oracle(ANT), Oracle(ANT)
    s0 = perturbed_states(ref, 1, seed=5, height=[0.5], qvel_std=0.3, qpos_std=0.05)[0]
    rng = np.random.default_rng(6)
    for o in (ref, ours):
        o.reset()
        o.set_state(s0)
    lib.mj_forward(ours.m, ours.d)
    ref.forward()
    assert np.array_equal(np.array(ref.dfield("qacc")), np.array(ours.dfield("qacc")))
    for t in range(40):
        c = rng.uniform(-1, 1, ref.size("nu"))
        ref.dfield("ctrl")[:] = c
        ours.dfield("ctrl")[:] = c
        ref.step()
        if t % 2:
            lib.mj_step(ours.m, ours.d)
        else:                                   # the split form: controls may change between the halves
            lib.mj_step1(ours.m, ours.d)
            lib.mj_step2(ours.m, ours.d)
        assert np.array_equal(ref.get_state(), ours.get_state()), t
        assert np.array_equal(np.array(ref.dfield("qfrc_constraint")), np.array(ours.dfield("qfrc_constraint")))
    lib.mj_forwardSkip(ours.m, ours.d, 1, 1)
    ref.forward()
    assert np.array_equal(np.array(ref.dfield("qacc")), np.array(ours.dfield("qacc")))
    lib.mjb_forget_model(ours.m)


def test_rollout_models_entry():
    """mjb_rollout_models: _unsafe_rollout's argument list with one mjModel pointer per environment"""
    import ctypes as C
    L = hostemu_lib()
    L.mjb_rollout_models.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_uint] + [C.c_void_p] * 5 + [C.c_int]
    m = mb.Model(ANT, library=L)
    nenv, nstep = 3, 5
    b = mb.Batch(m, nenv)
    s0 = b.get_state()
    s0[:, 3] = [0.4, 0.6, 0.8]
    ctrl = np.random.default_rng(1).uniform(-1, 1, (nenv, nstep, m.size("nu")))
    want = b.rollout(s0, ctrl)
    models = (C.c_void_p * nenv)(*([m.ptr] * nenv))
    got = np.zeros_like(want)
    rc = L.mjb_rollout_models(models, nenv, nstep, mb.STATE_CTRL, s0.ctypes.data, None, ctrl.ctypes.data, got.ctypes.data, None, -1)
    assert rc == 0 and np.array_equal(got, want)
    other = mb.Model(ANT, library=L)
    models[1] = other.ptr
    assert L.mjb_rollout_models(models, nenv, nstep, mb.STATE_CTRL, s0.ctypes.data, None, ctrl.ctypes.data, got.ctypes.data, None, -1) != 0


@pytest.mark.parametrize("solver", [mb.SOLVER_NEWTON, mb.SOLVER_PGS])
def test_box_and_cylinder_colliders_bit_exact(solver):
    """plane-box, plane-cylinder, sphere-box, sphere-cylinder, capsule-box, box-box (engine_collision_primitive.c:
    101-258,345-423, engine_collision_box.c:35-1068) on models/boxes.xml: nine free bodies tumbling on a plane and on
    each other, contact lists and every field compared bit for bit at several instants, and the whole rollout"""
    path = os.path.join(ROOT, "models", "boxes.mjb")
    nenv, nstep = 6, 150
    m, b, o = make_pair(path, solver, library=hostemu_lib(), nenv=nenv, nconmax=64, njmax=300)
    rng = np.random.default_rng(3)
    o.reset()
    s0 = np.tile(o.get_state(), (nenv, 1))
    nq = o.size("nq")
    for e in range(nenv):
        for k in range(9):                      # random heights, orientations and spins of the free bodies
            s0[e, 1 + 7 * k + 2] += rng.uniform(-0.05, 0.3)
            s0[e, 1 + 7 * k + 3:1 + 7 * k + 7] = rng.normal(size=4)
        s0[e, 1 + nq:] = rng.normal(0, 1.5, o.size("nv"))
    ctrl = np.zeros((nenv, nstep, 0))
    out = b.rollout(s0, ctrl)
    ref, stats, _ = o.rollout(s0, ctrl, nthread=4)
    assert stats[:, 3].sum() == 0 and (b.warnings() == 0).all()
    assert np.array_equal(out, ref)
    types = set()
    for t in range(4, 150, 5):
        compare_forward(b, o, ref[:, t, :], np.zeros((nenv, 0)), rtol=0, exact=True, check_dual=(solver == mb.SOLVER_PGS))
        g1, g2, nc = b.field("con_geom1"), b.field("con_geom2"), b.field("ncon")[:, 0]
        gt = o.mfield("geom_type")
        for e in range(nenv):
            for c in range(nc[e]):
                types.add((int(gt[g1[e, c]]), int(gt[g2[e, c]])))
    assert {(0, 5), (0, 6), (2, 5), (2, 6), (3, 6), (6, 6)} <= types, types      # every new collider produced contacts


@pytest.mark.parametrize("solver", [mb.SOLVER_NEWTON, mb.SOLVER_PGS])
def test_predefined_contact_pairs_bit_exact(solver):
    """<contact><pair> (engine_collision_driver.c:651-662,820-826,2046-2056): pairs merged by signature into the
    body-pair walk, their own condim / friction / solref / solimp / margin, the duplicated dynamic pair dropped, a
    pair the collision masks would not select - models/ant_pairs.xml"""
    path = os.path.join(ROOT, "models", "ant_pairs.mjb")
    nenv, nstep = 6, 120
    m, b, o = make_pair(path, solver, library=hostemu_lib(), nenv=nenv, nconmax=48, njmax=200)
    s0 = perturbed_states(o, nenv, seed=14, height=[0.3, 0.45, 0.7], qvel_std=0.8, qpos_std=0.15)
    ctrl = np.random.default_rng(15).uniform(-1, 1, (nenv, nstep, o.size("nu")))
    out = b.rollout(s0, ctrl)
    ref, stats, _ = o.rollout(s0, ctrl, nthread=4)
    assert stats[:, 3].sum() == 0 and (b.warnings() == 0).all()
    assert np.array_equal(out, ref)
    dims = set()
    for t in range(4, nstep, 10):
        compare_forward(b, o, ref[:, t, :], ctrl[:, t, :], rtol=0, exact=True, check_dual=(solver == mb.SOLVER_PGS))
        nc, cd = b.field("ncon")[:, 0], b.field("con_dim")
        for e in range(nenv):
            dims |= set(int(x) for x in cd[e, :nc[e]])
    assert {1, 3, 4} <= dims, dims      # the pairs' own condim values reached the contacts


@pytest.mark.parametrize("solver", [mb.SOLVER_NEWTON, mb.SOLVER_PGS])
def test_refsite_and_slidercrank_transmissions_bit_exact(solver):
    """site transmissions with a reference site (translational / rotational gear, common ancestral dofs cleared) and
    slider-crank transmissions (engine_core_smooth.c:1396-1465,1595-1700) - models/ant_trn.xml"""
    path = os.path.join(ROOT, "models", "ant_trn.mjb")
    nenv, nstep = 5, 100
    m, b, o = make_pair(path, solver, library=hostemu_lib(), nenv=nenv)
    s0 = perturbed_states(o, nenv, seed=14, height=[0.4, 0.55, 0.8], qvel_std=0.8, qpos_std=0.15)
    ctrl = np.random.default_rng(15).uniform(-1, 1, (nenv, nstep, o.size("nu")))
    compare_forward(b, o, s0, ctrl[:, 0], rtol=0, exact=True, check_dual=(solver == mb.SOLVER_PGS))
    out = b.rollout(s0, ctrl)
    ref, stats, _ = o.rollout(s0, ctrl, nthread=4)
    assert stats[:, 3].sum() == 0 and (b.warnings() == 0).all()
    assert np.array_equal(out, ref)
    for t in (20, 60, 99):
        compare_forward(b, o, ref[:, t, :], ctrl[:, t, :], rtol=0, exact=True, check_dual=(solver == mb.SOLVER_PGS))
        for e in range(nenv):     # lengths of the new transmissions
            o.reset(); o.set_state(ref[e, t]); o.dfield("ctrl")[:] = ctrl[e, t]; o.forward()
            assert np.array_equal(b.field("actuator_length")[e], np.array(o.dfield("actuator_length")))
            assert np.array_equal(b.field("actuator_velocity")[e], np.array(o.dfield("actuator_velocity")))


@pytest.mark.parametrize("model,solver", [(HUMANOID, mb.SOLVER_PGS), (HUMANOID, mb.SOLVER_NEWTON), (ANT, mb.SOLVER_CG),
                                          ("ant_frictionloss", mb.SOLVER_NEWTON), ("ant_condim", mb.SOLVER_PGS),
                                          ("ant_balls", mb.SOLVER_NEWTON)])
def test_noslip_post_solver_bit_exact(model, solver):
    """opt.noslip_iterations > 0 (solNoSlip, engine_solver.c:767-957): after the main solver - dual or primal - the
    friction rows are re-solved with the regulariser taken out of AR (dry-friction rows one by one, pyramidal contacts
    one pair of opposing edges at a time), then mj_dualFinish maps the forces back for every solver; the dual
    projection (efc_AR) is therefore also built for Newton / CG.  Monolithic and per island (ant_balls: several trees).
    Forward fields (forces, states, iteration counts incl. the noslip sweeps) and rollouts are bit-identical."""
    path = model if model.endswith(".mjb") else os.path.join(ROOT, "models", model + ".mjb")
    nenv, nstep = 6, 60
    m, b, o = make_pair(path, solver, library=hostemu_lib(), nenv=nenv, nconmax=48, njmax=220, noslip_iterations=4)
    assert o.opt("noslip_iterations") == 4
    s0 = perturbed_states(o, nenv, seed=31, height=[0.25, 0.4, 0.6], qvel_std=0.6, qpos_std=0.1)
    ctrl = np.random.default_rng(32).uniform(-1, 1, (nenv, nstep, o.size("nu")))
    compare_forward(b, o, s0, ctrl[:, 0], rtol=0, exact=True, check_dual=True)
    out = b.rollout(s0, ctrl)
    ref, stats, _ = o.rollout(s0, ctrl, nthread=4)
    assert stats[:, 0].sum() > 0 and stats[:, 3].sum() == 0
    assert np.array_equal(out, ref)
    # the post-solver changes the trajectory (otherwise this test would not see it)
    m2, b2, o2 = make_pair(path, solver, library=hostemu_lib(), nenv=nenv, nconmax=48, njmax=220)
    assert not np.array_equal(b2.rollout(s0, ctrl), out)


@pytest.mark.parametrize("solver", [mb.SOLVER_NEWTON, mb.SOLVER_PGS])
def test_contact_override_bit_exact(solver):
    """mjENBL_OVERRIDE (mj_assignMargin / Ref / Imp / Friction, engine_core_constraint.c:176-217): margin, solref, solimp
    and friction of every contact - dynamic geom pairs and predefined pairs - come from mjOption's o_* members; a pair's
    solreffriction is accepted under pyramidal cones (only elliptic friction rows would read it) -
    models/ant_override.xml"""
    path = os.path.join(ROOT, "models", "ant_override.mjb")
    nenv, nstep = 6, 100
    m, b, o = make_pair(path, solver, library=hostemu_lib(), nenv=nenv, nconmax=48, njmax=220)
    s0 = perturbed_states(o, nenv, seed=14, height=[0.35, 0.5, 0.75], qvel_std=0.8, qpos_std=0.05)
    ctrl = np.random.default_rng(15).uniform(-1, 1, (nenv, nstep, o.size("nu")))
    compare_forward(b, o, s0, ctrl[:, 0], rtol=0, exact=True, check_dual=(solver == mb.SOLVER_PGS))
    out = b.rollout(s0, ctrl)
    ref, stats, _ = o.rollout(s0, ctrl, nthread=4)
    assert stats[:, 0].sum() > 0 and stats[:, 3].sum() == 0
    assert np.array_equal(out, ref)
    e = int(np.argmax(b.field("ncon")[:, 0]))
    assert b.field("ncon")[e, 0] > 0 and np.allclose(b.field("con_friction")[e][:5], [0.7, 0.7, 0.01, 0.0002, 0.0002])
    assert np.allclose(b.field("con_solref")[e][:2], [0.015, 0.8])
