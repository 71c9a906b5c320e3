"""development probe: per-environment (nefc, solver_niter) of the bench workload after settling, saved for offline
analysis of the PGS critical path"""
import sys
import numpy as np
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import torch
import mujoco_b200 as mb
m = mb.Model('/root/repo/models/humanoid.mjb'); m.set_option('solver', 0)
nenv = 4096
b = mb.Batch(m, nenv)
nu, stride = m.size('nu'), b.env_stride()
stream = torch.cuda.ExternalStream(b.stream())
g = torch.Generator(device='cuda'); g.manual_seed(0)
out = {}
b.reset()
for rep in range(6):
    n = 300 if rep == 0 else 7
    c = (torch.rand((n, nu, stride), generator=g, device='cuda', dtype=torch.float64) * 2 - 1).contiguous()
    torch.cuda.synchronize(); b.rollout_device(n, c.data_ptr(), 0); stream.synchronize()
    out['nefc%d' % rep] = b.field('nefc')[:, 0].copy(); out['niter%d' % rep] = b.field('solver_niter')[:, 0].copy()
np.savez('/root/repo/gpurun_out/work.npz', **out)
print('saved')
