import sys, time, numpy as np
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import mujoco_b200 as mb
m = mb.Model('/root/repo/models/humanoid.mjb'); m.set_option('solver',0)
nenv=4096
for flags in [0, 1]:
    m.set_option('disableflags', flags)
    b = mb.Batch(m, nenv)
    b.reset()
    rng=np.random.default_rng(0)
    b.set_field('ctrl', rng.uniform(-1,1,(nenv,21)))
    b.step(300)
    for rep in range(3):
        ts=[]
        for s in range(4):
            t0=time.perf_counter(); b.run_stages(s,s); ts.append((time.perf_counter()-t0)*1e3)
        print('disable',flags,'stage wall ms', ['%.3f'%t for t in ts])
    t0=time.perf_counter(); b.step(50); print('50 fused steps: %.3f ms/step'%((time.perf_counter()-t0)*1e3/50))
    ne=b.field('nefc')[:,0]; it=b.field('solver_niter')[:,0]
    print('nefc mean/max', ne.mean(), ne.max(), 'iter mean/max', it.mean(), it.max(), 'sum iter*nefc^2', (it*ne*ne).mean())
