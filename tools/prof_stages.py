"""development probe: cycles per sub-stage of the split step's two halves (library built with -DMJB_STAGE_PROF)"""
import ctypes, sys
import numpy as np
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import torch
import mujoco_b200 as mb
cdll = ctypes.CDLL('/root/repo/prof_build/libmjb200_prof.so')
lib = mb._bind(cdll)
m = mb.Model('/root/repo/models/humanoid.mjb', library=lib); m.set_option('solver', 0)
nenv = 4096
b = mb.Batch(m, nenv)
nu, stride = m.size('nu'), b.env_stride()
stream = torch.cuda.ExternalStream(b.stream())
g = torch.Generator(device='cuda'); g.manual_seed(0)
b.reset()
names = ['kinematics+com+tendon', 'makeM+factor', 'collision', 'make_constraint', 'project (Y, AR)', 'transmission',
         'fwd_velocity', 'actuation+acceleration', 'constraint_begin', 'dual_finish', 'euler', 'project: half solve (Y)', 'project: AR']
for rep in range(3):
    n = 300 if rep == 0 else 10
    c = (torch.rand((n, nu, stride), generator=g, device='cuda', dtype=torch.float64) * 2 - 1).contiguous()
    torch.cuda.synchronize(); b.rollout_device(n, c.data_ptr(), 0); stream.synchronize()
    p1 = np.zeros(64, dtype=np.uint64); p2 = np.zeros(64, dtype=np.uint64)
    cdll.mjb_debug_stage_prof_p1(p1.ctypes.data_as(ctypes.c_void_p)); cdll.mjb_debug_stage_prof_p2(p2.ctypes.data_as(ctypes.c_void_p))
    if rep == 0: continue
    tot = p1 + p2
    print('sub-stage                      mean cycles/env-step    max cycles (any env, any step)')
    for i, nm in enumerate(names):
        print('%-30s %12.0f %18d' % (nm, float(tot[2 * i]) / (nenv * n), int(max(p1[2 * i + 1], p2[2 * i + 1]))))
