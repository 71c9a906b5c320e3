"""shared helpers for the parity tests (TEST INFRASTRUCTURE)"""
import ctypes as C
import os

import numpy as np

import mujoco_b200 as mb
from oracle_util import Oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOSTEMU = os.path.join(ROOT, "tests", "hostemu", "libmjb_hostemu.so")
HUMANOID = os.path.join(ROOT, "models", "humanoid.mjb")
ANT = os.path.join(ROOT, "models", "ant.mjb")

_hostemu = None


def hostemu_lib():
    """the kernel source compiled for the host; tests only, never loaded by the package"""
    global _hostemu
    if _hostemu is None:
        _hostemu = mb._bind(C.CDLL(HOSTEMU))
    return _hostemu


def make_pair(path, solver, library=None, nenv=4, nconmax=48, njmax=128, warp_per_env=True, **opts):
    """(product/hostemu model+batch, oracle) with identical option overrides"""
    m = mb.Model(path, library=library)
    o = Oracle(path)
    m.set_option("solver", solver)
    o.set_opt("solver", solver)
    for k, v in opts.items():
        m.set_option(k, v)
        o.set_opt(k, v)
    b = mb.Batch(m, nenv, nconmax=nconmax, njmax=njmax, warp_per_env=warp_per_env)
    return m, b, o


def perturbed_states(o, nenv, seed, height=None, qvel_std=1.0, qpos_std=0.3):
    """seeded random states around qpos0 (quaternions left unnormalised on purpose: the engine normalises)"""
    rng = np.random.default_rng(seed)
    o.reset()
    base = o.get_state()
    nq, nv = o.size("nq"), o.size("nv")
    out = np.tile(base, (nenv, 1))
    for e in range(nenv):
        out[e, 1:1 + nq] += rng.normal(0, qpos_std, nq)
        if height is not None:
            out[e, 3] = height[e % len(height)]
        out[e, 1 + nq:1 + nq + nv] = rng.normal(0, qvel_std, nv)
    return out


FIELDS_POS = ["xpos", "xquat", "xmat", "xipos", "ximat", "xanchor", "xaxis", "geom_xpos", "geom_xmat",
              "subtree_com", "cinert", "cdof", "crb", "M", "qLD", "qLDiagInv", "ten_length", "ten_J",
              "actuator_length"]
FIELDS_VEL = ["ten_velocity", "actuator_velocity", "cvel", "cdof_dot", "qfrc_spring", "qfrc_damper",
              "qfrc_passive", "qfrc_bias", "actuator_force", "qfrc_actuator", "qfrc_smooth", "qacc_smooth"]
FIELDS_EFC = ["efc_pos", "efc_margin", "efc_diagA", "efc_D", "efc_R", "efc_vel", "efc_aref", "efc_b", "efc_force"]


def compare_forward(b, o, states, ctrl, rtol, exact=False, check_dual=True):
    """run mj_forward per env on the oracle and compare every hot-path field; returns worst rel err"""
    b.set_state(states)
    b.set_field("ctrl", ctrl)
    b.set_field("qacc_warmstart", 0.0)     # the oracle side starts from mj_resetData
    b.forward()
    got = {f: b.field(f) for f in FIELDS_POS + FIELDS_VEL + FIELDS_EFC +
           ["efc_J", "efc_KBIP", "efc_AR", "efc_Y", "qacc", "qfrc_constraint", "con_dist", "con_pos", "con_frame",
            "con_includemargin", "con_friction", "con_solref", "con_solimp"]}
    gi = {f: b.field(f) for f in ["ncon", "nefc", "efc_type", "efc_id", "con_geom1", "con_geom2", "con_dim", "con_exclude",
                                  "con_efcadr", "solver_niter"]}
    worst = 0.0
    nv = o.size("nv")
    for e in range(states.shape[0]):
        o.reset()
        o.set_state(states[e])
        o.dfield("ctrl")[:] = ctrl[e]
        o.forward()
        ncon, nefc = int(o.scalar("ncon")), int(o.scalar("nefc"))
        assert gi["ncon"][e, 0] == ncon, (e, gi["ncon"][e, 0], ncon)
        assert gi["nefc"][e, 0] == nefc, (e, gi["nefc"][e, 0], nefc)
        assert np.array_equal(gi["efc_type"][e, :nefc], np.array(o.dfield("efc_type"))[:nefc])
        assert np.array_equal(gi["efc_id"][e, :nefc], np.array(o.dfield("efc_id"))[:nefc])
        nit = np.array(o.dfield("solver_niter"))
        assert np.array_equal(gi["solver_niter"][e, :len(nit)], nit), (e, gi["solver_niter"][e], nit)   # per island

        def chk(name, ref, n=None):
            nonlocal worst
            a = got[name][e]
            r = np.asarray(ref, dtype=np.float64).reshape(-1)
            if n is not None:
                a, r = a[:n], r[:n]
            assert a.shape == r.shape, (name, a.shape, r.shape)
            if exact:
                assert np.array_equal(a, r), (name, e, np.abs(a - r).max())
            else:
                scale = max(1.0, np.abs(r).max() if r.size else 1.0)
                err = (np.abs(a - r).max() / scale) if r.size else 0.0
                worst = max(worst, err)
                assert err <= rtol, (name, e, err)

        # the contact list (north_star: bit-exact contact counts / INDICES): geom pair, dimension, exclude flag and
        # constraint address of every contact are exact; distance / position / frame / mixed parameters at 1e-12
        con = o.contacts(max(ncon, 1))
        assert np.array_equal(gi["con_geom1"][e, :ncon], con["geom"][:, 0]), (e, gi["con_geom1"][e, :ncon], con["geom"][:, 0])
        assert np.array_equal(gi["con_geom2"][e, :ncon], con["geom"][:, 1]), (e, gi["con_geom2"][e, :ncon], con["geom"][:, 1])
        assert np.array_equal(gi["con_dim"][e, :ncon], con["dim"])
        assert np.array_equal(gi["con_exclude"][e, :ncon], con["exclude"])
        assert np.array_equal(gi["con_efcadr"][e, :ncon], con["efc_address"]), (e, gi["con_efcadr"][e, :ncon], con["efc_address"])
        for name, key, w in [("con_dist", "dist", 1), ("con_pos", "pos", 3), ("con_frame", "frame", 9),
                             ("con_includemargin", "includemargin", 1), ("con_friction", "friction", 5),
                             ("con_solref", "solref", 2), ("con_solimp", "solimp", 5)]:
            a, r = got[name][e][:ncon * w], con[key].reshape(-1)
            if a.size:
                err = np.abs(a - r).max() / max(1.0, np.abs(r).max())
                assert err <= 1e-12, (name, e, err)
        for f in FIELDS_POS + FIELDS_VEL:
            chk(f, o.dfield(f))
        for f in FIELDS_EFC:
            chk(f, o.dfield(f), nefc)
        chk("efc_KBIP", o.dfield("efc_KBIP"), 4 * nefc)
        chk("efc_J", o.dfield("efc_J"), nefc * nv)
        if check_dual and nefc:
            chk("efc_AR", o.dfield("efc_AR"), nefc * nefc)
            chk("efc_Y", o.dfield("efc_Y"), nefc * nv)
        chk("qacc", o.dfield("qacc"))
        chk("qfrc_constraint", o.dfield("qfrc_constraint"))
    return worst
