"""Throughput of the fused step on any supported BASELINE config (parity-test configs are not bench
lines; this is a development probe).  usage: bench_config.py <model.mjb> <solver 0|2> <nenv> [nstep] [settle]"""
import sys, time
import numpy as np
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import torch
import mujoco_b200 as mb

path, solver, nenv = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
nstep = int(sys.argv[4]) if len(sys.argv) > 4 else 200
settle = int(sys.argv[5]) if len(sys.argv) > 5 else 300
import os, ctypes
variant = os.environ.get('MJB_VARIANT_LIB')   # development only: A/B a differently-built library
m = mb.Model(path, library=mb._bind(ctypes.CDLL(variant)) if variant else None); m.set_option('solver', solver)
for kv in sys.argv[6:]:
    k, v = kv.split('='); m.set_option(k, float(v))
b = mb.Batch(m, nenv, nconmax=int(os.environ.get('MJB_NCONMAX', '0')), njmax=int(os.environ.get('MJB_NJMAX', '0')))   # 0 = library defaults
nu, stride = m.size('nu'), b.env_stride()
stream = torch.cuda.ExternalStream(b.stream())
g = torch.Generator(device='cuda'); g.manual_seed(0)
def ctrl(n): return (torch.rand((n, nu, stride), generator=g, device='cuda', dtype=torch.float64) * 2 - 1).contiguous()
b.reset()
c = ctrl(settle); torch.cuda.synchronize(); b.rollout_device(settle, c.data_ptr(), 0); stream.synchronize()
c = ctrl(nstep); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(stream); b.rollout_device(nstep, c.data_ptr(), 0); e1.record(stream); stream.synchronize()
ms = e0.elapsed_time(e1)
ne = b.field('nefc')[:, 0]; it = b.field('solver_niter')[:, 0]; nc = b.field('ncon')[:, 0]
print('%s solver=%d nenv=%d: %.3f ms/step, %.0f env-steps/s | ncon %.2f nefc mean %.2f max %d iter mean %.2f max %d warnings %d' %
      (path.split('/')[-1], solver, nenv, ms / nstep, nenv * nstep / ms * 1e3, nc.mean(), ne.mean(), ne.max(), it.mean(), it.max(), int(b.warnings().sum())))
