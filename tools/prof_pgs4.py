"""development probe: per-warp cycle counters of k_pgs4 (library built with -DMJB_PGS4_PROF)"""
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
for rep in range(3):
    n = 300 if rep == 0 else 5
    c = (torch.rand((n, nu, stride), generator=g, device='cuda', dtype=torch.float64) * 2 - 1).contiguous()
    torch.cuda.synchronize(); b.rollout_device(n, c.data_ptr(), 0); stream.synchronize()
    prof = np.zeros((nenv // 8, 8), dtype=np.int64)
    rc = cdll.mjb_debug_pgs4_prof(prof.ctypes.data_as(ctypes.c_void_p), nenv // 8)
    ne = b.field('nefc')[:, 0].reshape(-1, 8); it = b.field('solver_niter')[:, 0].reshape(-1, 8)
    tot, cyc, rows, sw, mom, prime, dce, pro = prof.T
    prime = prime - mom; dce = dce - cyc
    i = np.argsort(tot)[::-1][:6]
    print('rc', rc, 'max total cycles %d (%.0f us @1.9GHz) | cycles/row (all warps) %.1f | sweep overhead/sweep %.0f' % (
        tot.max(), tot.max() / 1.9e3, cyc.sum() / max(rows.sum(), 1), (tot - cyc).sum() / max(sw.sum(), 1)))
    for j in i:
        print('  warp %4d total %8d rowloop %8d rows %5d sweeps %3d -> %.0f cyc/row, %.0f other/sweep | nefc max %d | per sweep: momentum %.0f priming %.0f dce+end %.0f | prologue %d' % (
            j, tot[j], cyc[j], rows[j], sw[j], cyc[j] / max(rows[j], 1), (tot[j] - cyc[j]) / max(sw[j], 1), ne[j].max(), mom[j] / sw[j], prime[j] / sw[j], dce[j] / sw[j], pro[j]))
    for cls, lo, hi in (('NQ4', 1, 19), ('NQ8', 20, 35)):
        msk = (ne.max(axis=1) >= lo) & (ne.max(axis=1) <= hi)
        if msk.any(): print('  class %s: %d warps, %.0f cyc/row' % (cls, msk.sum(), cyc[msk].sum() / max(rows[msk].sum(), 1)))
