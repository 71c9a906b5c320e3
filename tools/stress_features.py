"""Randomised parity sweep of the CUDA path against the oracle over every feature model (development / evidence
tool; the assertions live in tests/).  For each model x solver: seeded random states and controls, one rollout,
per-step relative error against the reference engine; writes gpurun_out/r02_stress.json.
usage: python tools/stress_features.py [nenv] [nstep] [seeds]"""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import mujoco_b200 as mb
from mjb_util import make_pair, perturbed_states

nenv = int(sys.argv[1]) if len(sys.argv) > 1 else 32
nstep = int(sys.argv[2]) if len(sys.argv) > 2 else 80
seeds = int(sys.argv[3]) if len(sys.argv) > 3 else 2
MODELS = ["humanoid", "ant", "ant_frictionloss", "ant_balls", "ant_sensors", "ant_servo", "ant_equality", "ant_connect",
          "ant_weld", "ant_condim", "ant_fluid", "ant_touch", "ant_mocap", "ant_act", "ant_pairs", "ant_trn", "ant_override", "boxes"]
rows = []
CASES = [(n, {}) for n in MODELS] + [(n, {"noslip_iterations": 3}) for n in ("humanoid", "ant_condim", "ant_balls", "ant_frictionloss")]
for name, extra in CASES:
    path = os.path.join(ROOT, "models", name + ".mjb")
    for solver, sname in ((mb.SOLVER_PGS, "pgs"), (mb.SOLVER_NEWTON, "newton"), (mb.SOLVER_CG, "cg")):
        worst30 = worst = single = 0.0
        nwarn = 0
        t0 = time.time()
        try:
            for seed in range(seeds):
                m, b, o = make_pair(path, solver, nenv=nenv, nconmax=96, njmax=400, **extra)
                nq, nv, nu = o.size("nq"), o.size("nv"), o.size("nu")
                rng = np.random.default_rng(1000 * seed + 7)
                if name == "boxes":
                    o.reset(); s0 = np.tile(o.get_state(), (nenv, 1))
                    for e in range(nenv):
                        for k in range(9):
                            s0[e, 1 + 7 * k + 2] += rng.uniform(-0.05, 0.3)
                            s0[e, 1 + 7 * k + 3:1 + 7 * k + 7] = rng.normal(size=4)
                        s0[e, 1 + nq:] = rng.normal(0, 1.5, nv)
                else:
                    hs = [0.3, 0.45, 0.7, 1.0] if name != "humanoid" else [0.3, 0.6, 1.0, 1.3]
                    s0 = perturbed_states(o, nenv, seed=seed + 50, height=hs, qvel_std=0.6, qpos_std=0.15)
                ctrl = rng.uniform(-1, 1, (nenv, nstep, nu))
                out = b.rollout(s0, ctrl)
                ref, stats, _ = o.rollout(s0, ctrl, nthread=os.cpu_count() or 1)
                ok = stats[:, 3] == 0            # environments the reference stepped without a warning
                nwarn += int((~ok).sum())
                scale = np.maximum(1.0, np.abs(ref).max(axis=(0, 1)))
                err = (np.abs(out - ref) / scale)[ok].max(axis=(0, 2)) if ok.any() else np.zeros(nstep)
                worst30, worst = max(worst30, float(err[:30].max())), max(worst, float(err.max()))
                # ONE step from the reference's own state (both sides re-synchronised, warm start cleared): separates what a
                # single step contributes from what the dynamics amplify in a free-running trajectory
                if seed == 0:
                    for t in range(5, nstep - 1, max(1, (nstep - 6) // 6)):
                        b.set_state(ref[:, t, :]); b.set_field("qacc_warmstart", 0.0)
                        if nu: b.set_field("ctrl", ctrl[:, t + 1, :])
                        b.step(1)
                        got = b.get_state()
                        for e in range(0, nenv, 4):
                            if not ok[e]: continue
                            o.reset(); o.set_state(ref[e, t])
                            if nu: o.dfield("ctrl")[:] = ctrl[e, t + 1]
                            o.step()
                            r = o.get_state()
                            single = max(single, float(np.abs(got[e] - r).max() / max(1.0, np.abs(r).max())))
            rows.append({"model": name + ("+noslip" if extra else ""), "solver": sname, "nenv": nenv, "nstep": nstep, "seeds": seeds, "rel_err_30": worst30,
                         "rel_err_all": worst, "rel_err_single_step": single, "envs_with_reference_warnings": nwarn, "seconds": round(time.time() - t0, 2)})
        except mb.MjbError as ex:
            rows.append({"model": name + ("+noslip" if extra else ""), "solver": sname, "refused": str(ex)[:100]})
        print(rows[-1], flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump({"rows": rows}, open(os.path.join(ROOT, "gpurun_out", "r02_stress.json"), "w"), indent=1)
bad = [r for r in rows if "rel_err_30" in r and (r["rel_err_30"] > 1e-9 or r["rel_err_all"] > 1e-6)]
print("rows", len(rows), "outside (1e-9 @30 steps, 1e-6 overall):", len(bad))
for r in bad: print("  ", r)
