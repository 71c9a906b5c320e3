"""development probe: cycles per warp and phase inside the persistent rollout kernel (library built with
-DMJB_ROLLOUT_PROF by tools/build_prof.py)"""
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
W = 16
for rep in range(3):
    n = 300 if rep == 0 else 100
    c = (torch.rand((n, nu, stride), generator=g, device='cuda', dtype=torch.float64) * 2 - 1).contiguous()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    with torch.cuda.stream(stream):
        ev0.record(stream); b.rollout_device(n, c.data_ptr(), 0); ev1.record(stream)
    stream.synchronize()
    p = np.zeros((160, W, 8), dtype=np.uint64)
    cdll.mjb_debug_rollout_prof(p.ctypes.data_as(ctypes.c_void_p))
    if rep == 0: continue
    p = p[:147].astype(np.float64) / n      # cycles per step; CTAs with a full block of environments
    names = ['first half (own pair)', 'wait at barrier 1', 'PGS (own envs) / idle', 'wait at barrier 2', 'second half (own pair)']
    print('%.3f ms/step; cycles per step, mean over CTAs [min .. max over CTAs]:' % (ev0.elapsed_time(ev1) / n))
    tot = p[:, :, :5].sum(axis=2)
    print('  total per warp: mean %.0f  max-CTA %.0f  min-CTA %.0f' % (tot.mean(), tot.mean(axis=1).max(), tot.mean(axis=1).min()))
    for i, nm in enumerate(names):
        halves = p[:, :14, i]; pg = p[:, :4, i]
        print('  %-26s halves-warps mean %8.0f [%8.0f .. %8.0f]   PGS warps (0-3) mean %8.0f' %
              (nm, halves.mean(), halves.mean(axis=1).min(), halves.mean(axis=1).max(), pg.mean()))
    print('  slowest warp of a CTA in first half, mean over CTAs: %.0f ; second half: %.0f' % (p[:, :14, 0].max(axis=1).mean(), p[:, :14, 4].max(axis=1).mean()))
