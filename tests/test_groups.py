"""Grouped execution of multi-step calls (mjb_api.cc: env_groups): the batch is cut into contiguous env
groups that advance independently.  On the host emulation (no streams) this checks the OFFSET arithmetic
of every buffer layout involved — reference layout [env][step][n] and native layout [step][n][env] —
by comparing a grouped run with an ungrouped one, bit for bit.  The group count is read once per
process, hence the subprocesses."""
import os
import subprocess
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)

SCRIPT = r'''
import sys, numpy as np
sys.path.insert(0, %(root)r); sys.path.insert(0, %(here)r)
import mujoco_b200 as mb
from mjb_util import HUMANOID, hostemu_lib
nenv, nstep = 330, 3                      # not a multiple of the group size: last group is short
m = mb.Model(HUMANOID, library=hostemu_lib()); m.set_option("solver", mb.SOLVER_PGS)
b = mb.Batch(m, nenv, nconmax=64, njmax=200)   # generous caps: a warning would stop an env in rollout() only
rng = np.random.default_rng(3)
b.reset(); s0 = b.get_state()
s0[:, 3] = rng.uniform(0.3, 1.3, nenv); s0[:, 1 + 28:] = rng.normal(0, 0.3, (nenv, 27))
ctrl = rng.uniform(-1, 1, (nenv, nstep, 21))
out = b.rollout(s0, ctrl)                 # reference layouts
# native layouts: ctrl [nstep][nu][stride], state [nstep][nstate][stride]
stride = b.env_stride()
cn = np.zeros((nstep, 21, stride)); cn[:, :, :nenv] = ctrl.transpose(1, 2, 0)
sn = np.zeros((nstep, 56, stride))
b.set_state(s0); b.set_field("qacc_warmstart", np.zeros((nenv, 27)))
b.rollout_device(nstep, cn.ctypes.data, sn.ctypes.data)
b.set_state(s0); b.set_field("qacc_warmstart", np.zeros((nenv, 27)))
b.step(nstep)
assert b.warnings().sum() == 0
np.savez(sys.argv[1], out=out, native=sn[:, :, :nenv].transpose(2, 0, 1), stepped=b.get_state())
'''


def _run(tmp_path, groups):
    out = str(tmp_path / ("g%d.npz" % groups))
    env = dict(os.environ, MJB_GROUPS=str(groups))
    subprocess.run([sys.executable, "-c", SCRIPT % {"root": ROOT, "here": HERE}, out], check=True, env=env, timeout=600)
    return np.load(out)


@pytest.mark.timeout(900)
def test_grouped_equals_ungrouped(tmp_path):
    from mjb_util import HOSTEMU
    if not os.path.exists(HOSTEMU):
        pytest.skip("host emulation library not built")
    a, b = _run(tmp_path, 1), _run(tmp_path, 4)
    for k in ("out", "native", "stepped"):
        assert np.array_equal(a[k], b[k]), k
    # the reference-layout and native-layout entry points agree with each other as well
    # ("stepped" keeps the controls of the last step, so it is only compared across group counts)
    assert np.array_equal(a["out"], a["native"])
