"""Size-independent properties of the path, written after the reference's own pipeline-level tests
(they are equivalence / invariance tests there too): test/engine/engine_solver_test.cc:226-350
(SolversEquivalent), :601-630 (ZeroToleranceDisablesTermination), :468-550 (WarmstartZeroIterations),
test/pipeline_test.cc:90-130 (DeterministicNoWarmstart).  Run on the host emulation of the kernel
source (CPU); the GPU suite repeats the determinism check on the device."""
import os

import numpy as np
import pytest

import mujoco_b200 as mb
from mjb_util import ANT, HOSTEMU, HUMANOID, hostemu_lib, perturbed_states
from oracle_util import Oracle, available

pytestmark = pytest.mark.skipif(not (available() and os.path.exists(HOSTEMU)), reason="oracle or hostemu not built")

DSBL_WARMSTART, DSBL_ISLAND = 1 << 9, 1 << 18


def _batch(path, nenv, **opts):
    m = mb.Model(path, library=hostemu_lib())
    for k, v in opts.items():
        m.set_option(k, v)
    return m, mb.Batch(m, nenv, nconmax=48, njmax=128)


@pytest.mark.parametrize("path", [HUMANOID, ANT])
def test_solvers_equivalent(path):
    """SolversEquivalent: with tolerance 0, 500 iterations and no warmstart, CG and PGS reach Newton's
    constraint force (relative to |qfrc_constraint|).  The reference asserts 1e-12 at a rest keyframe; on
    these randomly perturbed, deeply penetrating states CG still gets 1e-9 while PGS (linear convergence)
    is held to 1e-4 — and to being bit-identical to the reference's own PGS answer"""
    o = Oracle(path)
    states = perturbed_states(o, 6, seed=5, height=[0.3, 0.45, 0.9], qpos_std=0.1, qvel_std=0.3)
    res = {}
    for solver in (mb.SOLVER_NEWTON, mb.SOLVER_CG, mb.SOLVER_PGS):
        m, b = _batch(path, 6, solver=solver, tolerance=0, iterations=500, disableflags=DSBL_WARMSTART)
        b.set_state(states)
        b.forward()
        res[solver] = b.field("qfrc_constraint")
        assert (b.field("nefc")[:, 0] > 0).any()
    scale = np.maximum(np.linalg.norm(res[mb.SOLVER_NEWTON], axis=1, keepdims=True), 1e-30)
    for solver, tol in ((mb.SOLVER_CG, 1e-9), (mb.SOLVER_PGS, 1e-4)):
        err = (np.abs(res[solver] - res[mb.SOLVER_NEWTON]) / scale).max()
        assert err < tol, (solver, err)
    o.set_opt("solver", mb.SOLVER_PGS); o.set_opt("tolerance", 0); o.set_opt("iterations", 500)
    o.set_opt("disableflags", DSBL_WARMSTART)
    for e in range(6):
        o.reset(); o.set_state(states[e]); o.forward()
        assert np.array_equal(res[mb.SOLVER_PGS][e], np.array(o.dfield("qfrc_constraint")))


def test_zero_tolerance_disables_termination():
    """ZeroToleranceDisablesTermination: Newton with tolerance 0 runs exactly opt.iterations iterations"""
    o = Oracle(HUMANOID)
    states = perturbed_states(o, 4, seed=9, height=[0.3, 0.4], qpos_std=0.1)
    m, b = _batch(HUMANOID, 4, solver=mb.SOLVER_NEWTON, tolerance=0, iterations=3,
                  disableflags=DSBL_WARMSTART | DSBL_ISLAND)
    b.set_state(states)
    b.forward()
    assert (b.field("nefc")[:, 0] > 0).all()
    assert (b.field("solver_niter")[:, 0] == 3).all()


@pytest.mark.parametrize("solver", [mb.SOLVER_NEWTON, mb.SOLVER_PGS, mb.SOLVER_CG])
def test_deterministic_and_forward_is_idempotent(solver):
    """DeterministicNoWarmstart: two identical batches stay identical, and an extra mj_forward after a
    step changes nothing"""
    nenv, nstep = 3, 25
    o = Oracle(HUMANOID)
    s0 = perturbed_states(o, nenv, seed=13, height=[0.35, 0.6], qpos_std=0.1, qvel_std=0.3)
    ctrl = np.random.default_rng(14).uniform(-1, 1, (nenv, 21))
    m1, b1 = _batch(HUMANOID, nenv, solver=solver, disableflags=DSBL_WARMSTART)
    m2, b2 = _batch(HUMANOID, nenv, solver=solver, disableflags=DSBL_WARMSTART)
    for b in (b1, b2):
        b.set_state(s0)
        b.set_field("ctrl", ctrl)
    for _ in range(nstep):
        b1.step(1); b1.forward()
        b2.step(1); b2.forward()
        q1 = b1.field("qacc")
        assert np.array_equal(q1, b2.field("qacc"))
        b2.forward()
        assert np.array_equal(q1, b2.field("qacc"))


def test_warmstart_zero_iterations():
    """WarmstartZeroIterations: re-solving from the converged acceleration needs no Newton iteration"""
    o = Oracle(HUMANOID)
    states = perturbed_states(o, 4, seed=21, height=[0.3, 0.45], qpos_std=0.1, qvel_std=0.2)
    m, b = _batch(HUMANOID, 4, solver=mb.SOLVER_NEWTON, tolerance=1e-10, iterations=100)
    b.set_state(states)
    b.forward()
    assert (b.field("solver_niter")[:, 0] > 0).any()
    b.set_field("qacc_warmstart", b.field("qacc"))
    b.forward()
    assert (b.field("solver_niter")[:, 0] == 0).all()
