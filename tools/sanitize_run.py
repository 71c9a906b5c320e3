"""small rollouts of every supported config for compute-sanitizer (memcheck / racecheck / initcheck)"""
import sys
import numpy as np
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import mujoco_b200 as mb
from mjb_util import ANT, HUMANOID
nenv, nstep = int(sys.argv[1]) if len(sys.argv) > 1 else 32, int(sys.argv[2]) if len(sys.argv) > 2 else 40
for path, solver in [(HUMANOID, 0), (HUMANOID, 2), (ANT, 2), (ANT, 0)]:
    m = mb.Model(path); m.set_option('solver', solver)
    b = mb.Batch(m, nenv)
    b.reset()
    s0 = b.get_state()
    rng = np.random.default_rng(1)
    s0[:, 3] = rng.uniform(0.2, 1.0, nenv)          # drop heights: contacts early
    s0[:, 1 + m.size('nq'):] = rng.normal(0, 0.5, (nenv, m.size('nv')))
    ctrl = rng.uniform(-1, 1, (nenv, nstep, m.size('nu')))
    out = b.rollout(s0, ctrl)
    print(path.split('/')[-1], solver, 'finite', bool(np.isfinite(out).all()), 'nefc max', int(b.field('nefc').max()), 'warn', int(b.warnings().sum()))
