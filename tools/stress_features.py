"""CPU stress: every feature model x solver x integrator, random states and controls, kernel source (host build)
against the reference engine, bit for bit.  usage: python tools/stress_features.py"""
import os, sys, numpy as np, time
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import mujoco_b200 as mb
from mjb_util import ROOT, make_pair, perturbed_states, hostemu_lib
models = ["ant_equality","ant_connect","ant_weld","ant_mocap","ant_act","ant_act_nomuscle","ant_fluid","ant_condim","ant_sensors","ant_servo","ant_balls","ant_frictionloss"]
bad = 0
for name in models:
    path = os.path.join(ROOT,"models",name+".mjb")
    for solver in (0,1,2):
        for integ in (0,1,3):
            if integ==3 and name in ("ant_act","ant_fluid","ant_balls","ant_mocap","ant_weld","ant_connect","ant_condim"): continue
            try:
                m,b,o = make_pair(path, solver, library=hostemu_lib(), nenv=16, nconmax=64, njmax=260, integrator=integ)
            except Exception as ex:
                print(name, solver, integ, "refused:", str(ex)[:80]); continue
            nq,nv,na=o.size("nq"),o.size("nv"),o.size("na")
            for seed in (101,202):
                s0 = perturbed_states(o, 16, seed=seed, height=[0.3,0.45,0.6,0.9], qvel_std=1.0, qpos_std=0.12)
                if na: s0[:,1+nq+nv:] = np.random.default_rng(seed).uniform(-0.5,1,(16,na))
                ctrl = np.random.default_rng(seed+1).uniform(-1,1,(16,150,o.size("nu")))
                out = b.rollout(s0, ctrl)
                ref, stats, _ = o.rollout(s0, ctrl, nthread=8)
                ok = np.array_equal(out, ref)
                if not ok:
                    bad += 1
                    d = np.argwhere(out != ref)
                    print("MISMATCH", name, solver, integ, seed, "first at env/step", d[0][:2], "warn", int(b.warnings().sum()), "refwarn", int(stats[:,3].sum()))
print("done, mismatches:", bad)
