#!/usr/bin/env python
"""bench.py — env-steps/sec of the batched mj_step hot path (BASELINE.json metric).

Workload (config.workload): BASELINE configs[1] — model/humanoid/humanoid.xml, 4096 environments per
GPU, PGS solver, Euler integrator, i.i.d. uniform random ctrl in [-1,1] (seeded), fp64.
A "step" is one mj_step of every environment of the batch.  Before timing, every environment is
rolled SETTLE steps (untimed, fixed, not part of --warmup) so the timed window is the contact-rich
steady state (humanoids on the floor), then W warm-up steps, then exactly K timed steps.

  value   : HBM-resident throughput — controls pre-generated on the device, states stay on the device
  e2e     : through the reference-facing C-ABI call with HOST buffers: per step, ctrl [nenv,nu] is
            copied from pinned host memory, the step runs, the FULLPHYSICS state [nenv,nstate] is
            copied back (mjb_step_host)
  roofline: for the dominant kernel (k_stage stage 2 = constraint solve unless the stage timing pass
            says otherwise): algorithmic bytes / launch over its CUDA-event time
  cpu_baseline / --impl reference: the UNMODIFIED reference engine (oracle/_ref/libmujoco_ref.so,
            compiled from /root/reference) stepping the same workload on the host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

SETTLE = int(os.environ.get("MJB_BENCH_SETTLE", 300))
# BASELINE.json configs: [1] is the configuration the metric is quoted on (the default); [2] is the other
# configuration this build runs at full size.  configs[3] (humanoid100) and [4] (cube) need colliders / the
# sparse Jacobian that mjb_check_model still refuses - `--config 3|4` says so instead of printing a number.
CONFIGS = {
    1: dict(model="humanoid.mjb", nenv=4096, solver=0, integrator=0, name="humanoid.xml", sname="PGS", iname="Euler"),
    2: dict(model="ant.mjb", nenv=65536, solver=2, integrator=0, name="ant.xml", sname="Newton", iname="Euler"),
}
NENV = 4096
MODEL = METRIC = WORKLOAD = None
CFG = None


def select_config(idx):
    global NENV, MODEL, METRIC, WORKLOAD, CFG
    CFG = CONFIGS[idx]
    NENV = int(os.environ.get("MJB_BENCH_NENV", CFG["nenv"]))
    MODEL = os.path.join(ROOT, "models", CFG["model"])
    METRIC = "env-steps/sec (whole box) %s batch %d/GPU, %s, %s, random ctrl" % (CFG["name"], NENV, CFG["sname"], CFG["iname"])
    WORKLOAD = "%s x%d envs/GPU, solver=%s, integrator=%s, ctrl~U[-1,1] (configs[%d])" % (
        CFG["name"], NENV, CFG["sname"], CFG["iname"], idx)


def config_dict(l2_note):
    """identical keys (and values) in both arms, so that the driver can compare the two lines' `config`"""
    return {"workload": WORKLOAD, "nenv_per_gpu": NENV, "settle_steps": SETTLE, "l2": l2_note}


def l2_note(nq, nv, nu):
    # every step touches the hot block of every environment; the batch is far larger than the 126 MB L2
    return "no flush: inputs larger than L2 (per-env blocks of the whole batch are touched every step)"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p)).get("hbm_gbs", 6650.0), "measured"
    return 6650.0, "fallback"


class ClockSampler(threading.Thread):
    def __init__(self, index=0):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], False

    def run(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                self.rows.append([x.strip() for x in out.strip().split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        sm, mx, reasons = [], 0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = max(mx, float(r[1]))
                for n, v in zip(names, r[2:6]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons)}


def reference_arm(args, rank, world):
    """the reference's own CPU mj_step (unmodified engine) on the same workload, all host threads"""
    if rank != 0:
        return
    from oracle_util import Oracle, available
    if not available():
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libmujoco_ref.so not built on this box"}))
        return
    o = Oracle(MODEL)
    o.set_opt("solver", CFG["solver"])
    o.set_opt("integrator", CFG["integrator"])
    nu, cores = o.size("nu"), os.cpu_count() or 1
    nthread = cores
    # each "step" = one mj_step of a bounded sample of the batch, sized to ~0.25 s per step
    sample = NENV
    rng = np.random.default_rng(0)
    o.reset()
    s0 = np.tile(o.get_state(), (sample, 1))
    # settle to the contact-rich regime like the GPU arm
    ctrl = rng.uniform(-1, 1, (sample, SETTLE, nu))
    st, _, _ = o.rollout(s0, ctrl, nthread=nthread, want_state=True)
    s0 = st[:, -1, :].copy()
    if args.warmup:
        ctrl = rng.uniform(-1, 1, (sample, args.warmup, nu))
        st, _, _ = o.rollout(s0, ctrl, nthread=nthread, want_state=True)
        s0 = st[:, -1, :].copy()
    ctrl = rng.uniform(-1, 1, (sample, args.steps, nu))
    _, stats, sec = o.rollout(s0, ctrl, nthread=nthread, want_state=False)
    value = sample * args.steps / sec
    line = {
        "metric": METRIC, "value": value, "unit": "env-steps/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * sec / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic", "impl": "reference",
        "config": config_dict(l2_note(0, 0, 0)),
        "cpu_baseline": {"value": value, "unit": "env-steps/s", "cores": nthread, "kind": "reference",
                         "sample": "%d of %d envs x %d steps, mjo_rollout (rollout.cc loop), %d worker threads created "
                                   "before the clock starts" % (sample, NENV, args.steps, nthread)},
        "e2e": {"value": value, "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "workload_stats": {"mean_ncon": float(stats[:, 0].mean() / args.steps),
                           "mean_nefc": float(stats[:, 1].mean() / args.steps)},
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=1, help="BASELINE.json configs index (1 = the metric's configuration)")
    args = ap.parse_args()
    if args.config not in CONFIGS:
        print(json.dumps({"impl": args.impl, "config": {"workload": "BASELINE configs[%d]" % args.config},
                          "unavailable": "mjb_check_model refuses this model (colliders / sparse Jacobian not built)"}))
        return
    select_config(args.config)
    rank = int(os.environ.get("RANK", 0))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))

    if args.impl == "reference":
        reference_arm(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    import mujoco_b200 as mb

    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    K, W = args.steps, max(args.warmup, 3)

    model = mb.Model(MODEL)
    model.set_option("solver", CFG["solver"])
    model.set_option("integrator", CFG["integrator"])
    nu, nq, nv = model.size("nu"), model.size("nq"), model.size("nv")
    batch = mb.Batch(model, NENV, device=local_rank)
    stride = batch.env_stride()
    stream = torch.cuda.ExternalStream(batch.stream(), device=local_rank)
    nstate = 1 + nq + nv

    gen = torch.Generator(device="cuda")
    gen.manual_seed(1000 + rank)

    def make_ctrl(n):   # native layout [n][nu][stride]
        return (torch.rand((n, nu, stride), generator=gen, device="cuda", dtype=torch.float64) * 2 - 1).contiguous()

    # ---- settle (untimed) + warm-up
    batch.reset()
    c = make_ctrl(SETTLE)
    torch.cuda.synchronize()
    batch.rollout_device(SETTLE, c.data_ptr(), 0)
    stream.synchronize()
    c = make_ctrl(W)
    torch.cuda.synchronize()
    batch.rollout_device(W, c.data_ptr(), 0)
    stream.synchronize()

    # ---- timed region: K steps, controls and states HBM-resident
    c = make_ctrl(K)
    torch.cuda.synchronize()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = batch.kernel_launches()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    d_last = torch.empty((1, nstate, stride), device="cuda", dtype=torch.float64)
    c_last = c[K - 1:].contiguous()
    gathered = torch.empty((world, NENV), device="cuda", dtype=torch.float64) if world > 1 else None
    torch.cuda.synchronize()
    ev0.record(stream)   # nothing but the launches of the K steps (and the boundary collective) is enqueued after this
    batch.rollout_device(K - 1, c.data_ptr(), 0)
    batch.rollout_device(1, c_last.data_ptr(), d_last.data_ptr())
    if world > 1:
        # boundary collective: ONE all-gather of per-env episode returns (final torso height),
        # enqueued on the batch stream right behind the last step
        with torch.cuda.stream(stream):
            ret = d_last[0, 3, :NENV].contiguous()
            dist.all_gather_into_tensor(gathered, ret)
    ev1.record(stream)
    stream.synchronize()
    torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1)
    launches = batch.kernel_launches() - launches0
    if world > 1:
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    sampler.stop_flag = True
    value = NENV * world * K / (ms * 1e-3)

    # workload statistics at the end of the timed window
    ncon = batch.field("ncon")[:, 0].astype(np.float64)
    nefc = batch.field("nefc")[:, 0].astype(np.float64)
    niter = batch.field("solver_niter")[:, 0].astype(np.float64)
    warn = int(batch.warnings().sum())

    # ---- per-launch timing pass: CUDA events on the launching stream around each launch of the step (mjb_step_profile)
    nprobe = 20
    names = ["position+velocity", "solve", "finish+integrate", "redo"]
    launch_ms = np.zeros(4)
    split = True
    torch.cuda.synchronize()
    try:
        for t_ in range(nprobe):
            launch_ms += batch.step_profile()
        launch_ms /= nprobe
    except RuntimeError:       # this configuration steps with one fused launch
        split = False
    # ---- roofline of the dominant kernel
    peak, peak_src = peaks()
    m_nefc, m_nefc2, m_ncon = float(nefc.mean()), float((nefc ** 2).mean()), float(ncon.mean())
    m_iter = float(niter.mean())
    # ALGORITHMIC bytes per env-step, SURVEY.md section 8(d) convention "B_mjdata" (DESIGN.md section 5):
    #   B_state + 8 * (hot-path mjData fields written once) + ncon*584 + nefc*(8*(nv+14)+12)
    #   + PGS: 8*nefc*nv (efc_Y) + 8*nefc^2 (efc_AR)
    # with the batch's own mean ncon / nefc / nefc^2 at the end of the timed window; the per-field counts come from
    # the batch layout (mjb_field_size), so the formula follows the model.
    fs = batch.field_size
    pos_fields = ["xpos", "xquat", "xmat", "xipos", "ximat", "xanchor", "xaxis", "geom_xpos", "geom_xmat", "subtree_com",
                  "cdof", "cinert", "ten_J", "ten_length", "actuator_length", "actuator_moment", "crb", "M", "qLD", "qLDiagInv"]
    vel_fields = ["ten_velocity", "actuator_velocity", "cvel", "cdof_dot", "qfrc_bias", "qfrc_spring", "qfrc_damper",
                  "qfrc_passive", "qH", "qHDiagInv"]
    acc_fields = ["actuator_force", "qfrc_actuator", "qfrc_smooth", "qfrc_constraint", "qacc_smooth", "qacc"]
    n_pos, n_vel, n_acc = (sum(fs(f) for f in grp) for grp in (pos_fields, vel_fields, acc_fields))
    b_state = 8 * ((1 + nq + nv + nu + nv) + (1 + nq + nv + nv))
    pgs = CFG["solver"] == 0
    b_pos = 8 * n_pos + m_ncon * 584 + m_nefc * (8 * (nv + 14) + 12) + (8 * m_nefc * nv + 8 * m_nefc2 if pgs else 0)
    b_mjdata = b_state + b_pos + 8 * n_vel + 8 * n_acc
    step_gbs = NENV * b_mjdata / (ms / K * 1e-3) / 1e9
    prof = {}
    pp = os.path.join(ROOT, "profiles", "r02_kernels.json")
    if os.path.exists(pp):
        prof = json.load(open(pp))
    if split:
        dom = int(np.argmax(launch_ms[:3]))
        # algorithmic bytes of each launch: the fields it reads once and writes once
        b_k = [b_state / 2 + b_pos + 8 * n_vel + 8 * (n_acc - 2 * nv) + 8 * 2 * m_nefc,           # first half (+ efc_b, warm-start force)
               8 * m_nefc2 + 8 * 3 * m_nefc + 12 * m_nefc,                                        # solve: AR, b / force / frictionloss in, force + state out
               b_state / 2 + 8 * 2 * nv + 8 * m_nefc * nv + 8 * m_nefc]                           # second half: J' f, qacc, state out
        kname = ["k_step_warp<PGS,16,lean,part 1> (position + velocity)", "k_pgs4 (PGS, 4 lanes per environment)",
                 "k_step_warp<PGS,16,lean,part 2> (finish + integrate)"][dom]
        kkey = ["part1", "pgs4", "part2"][dom]
        achieved = NENV * b_k[dom] / (launch_ms[dom] * 1e-3) / 1e9
        traffic = prof.get(kkey, {}).get("dram_bytes_per_launch")
        roof = {"bound": "hbm", "kernel": kname, "achieved": achieved, "peak": peak, "peak_source": peak_src, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "traffic_source": prof.get("source"),
                "algorithmic_bytes_per_launch": NENV * b_k[dom], "launch_ms": float(launch_ms[dom]),
                "fp64_pipe_frac": prof.get(kkey, {}).get("fp64_pipe_frac"), "issue_active": prof.get(kkey, {}).get("issue_active"),
                "note": "not bandwidth-bound: the solve is ONE dependent chain per environment (Gauss-Seidel recurrence) and the launch "
                        "lasts as long as the batch's longest chain; see DESIGN.md section 5",
                "launches_ms": dict(zip(names, [float(x) for x in launch_ms])),
                "launches_bytes_per_env": dict(zip(names[:3], [float(x) for x in b_k])),
                "step": {"algorithmic_bytes_per_env_step": b_mjdata, "achieved": step_gbs, "frac": step_gbs / peak}}
    else:
        roof = {"bound": "hbm", "kernel": "k_step_warp (fused mj_step, 1 launch/step)", "achieved": step_gbs, "peak": peak,
                "peak_source": peak_src, "unit": "GB/s", "frac": step_gbs / peak, "traffic": None,
                "algorithmic_bytes_per_launch": NENV * b_mjdata, "launch_ms": ms / K,
                "note": "not bandwidth-bound: dependent chains of the constraint solver and the tree passes (DESIGN.md section 5)"}

    # ---- e2e: per step H2D ctrl from pinned host, step, D2H state
    ke = min(K, 100)
    h_ctrl = torch.rand((ke, NENV, nu), dtype=torch.float64).mul_(2).sub_(1).pin_memory()
    h_state = torch.empty((NENV, nstate), dtype=torch.float64).pin_memory()
    for t_ in range(3):
        batch.step_host(h_ctrl[t_], h_state)
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for t_ in range(ke):
        batch.step_host(h_ctrl[t_], h_state)
    e2e_s = time.perf_counter() - t0
    if world > 1:
        tt = torch.tensor([e2e_s], device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e_s = float(tt.item())
    e2e = {"value": NENV * world * ke / e2e_s, "unit": "env-steps/s", "h2d_bytes_per_step": NENV * nu * 8,
           "d2h_bytes_per_step": NENV * nstate * 8, "steps": ke,
           "call": "mjb_step_host per step: ctrl [nenv,nu] host -> device, one mj_step, FULLPHYSICS state device -> host, synchronised"}
    # the reference's own batched call shape (rollout.cc): nstep steps per call, controls for all steps in, states
    # of all steps out, host buffers
    kr = min(K, 20)
    # page-locked arrays for the call's inputs and outputs (the reference API takes preallocated `state=` the same way)
    h_c = torch.from_numpy(np.random.default_rng(3).uniform(-1, 1, (NENV, kr, nu))).pin_memory().numpy()
    h_out = torch.empty((NENV, kr, nstate), dtype=torch.float64).pin_memory().numpy()
    s_now = batch.get_state()
    batch.rollout(s_now, h_c, state=h_out)
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    batch.rollout(s_now, h_c, state=h_out)
    r_s = time.perf_counter() - t0
    e2e["rollout_call"] = {"value": NENV * world * kr / r_s, "unit": "env-steps/s", "steps_per_call": kr,
                           "h2d_bytes_per_step": NENV * nu * 8, "d2h_bytes_per_step": NENV * nstate * 8,
                           "call": "mjb_rollout (the rollout.cc contract): control [nenv,nstep,nu] in, states [nenv,nstep,nstate] out, page-locked host arrays, initial state set per call"}

    # ---- CPU baseline (rank 0, N=1 only): reference engine on a bounded sample, all host cores
    cpu = None
    if rank == 0 and world == 1:
        try:
            from oracle_util import Oracle, available
            if available():
                o = Oracle(MODEL)
                o.set_opt("solver", CFG["solver"])
                o.set_opt("integrator", CFG["integrator"])
                cores = os.cpu_count() or 1
                sample, csteps = min(NENV, 8192), 100
                s_now = batch.get_state()[:sample]
                rng = np.random.default_rng(1)
                cctrl = rng.uniform(-1, 1, (sample, csteps, nu))
                _, cst, sec = o.rollout(s_now, cctrl, nthread=cores, want_state=False)
                cpu = {"value": sample * csteps / sec, "unit": "env-steps/s", "cores": cores, "kind": "reference",
                       "sample": "%d envs (GPU batch states at end of timed window) x %d steps, %d threads" % (sample, csteps, cores)}
        except Exception as ex:  # noqa: BLE001
            cpu = {"value": None, "unit": "env-steps/s", "cores": os.cpu_count(), "kind": "reference", "sample": "failed: %s" % ex}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "env-steps/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": config_dict(l2_note(nq, nv, nu)),
            "workload_stats": {"mean_ncon": m_ncon, "mean_nefc": m_nefc, "mean_solver_iter": m_iter, "warnings": warn,
                               "bytes_touched_per_step_mb": b_mjdata * NENV / 1e6},
            "roofline": roof, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches),
            "clocks": sampler.summary(),
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
