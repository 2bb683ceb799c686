#!/usr/bin/env python
"""bench.py -- shot-gradient throughput of the FWI hot path on B200.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's own op (oracle/_ref), rank 0 only

Workload (BASELINE.json configs[1], "C2"): Marmousi-sized 134x384 model (224x448 padded), 379 receivers,
2000 time steps, 30 shots PER GPU (weak scaling), forward + adjoint gradient as `fwi_op`'s gradient kernel
computes it (calc_id 1).  One "step" = one gradient evaluation of the rank's 30 shots.
  value  : shot-gradients/s of the whole job, inputs (model, stf, observed data) resident in HBM, timed with
           CUDA events on the launching stream, barrier + synchronize on both sides, max over ranks;
           for N > 1 every step ends with the NCCL all-reduce of [grad_lambda|grad_mu|grad_den|misfit].
  e2e    : same metric through the reference-facing C-ABI call with HOST buffers (fwi_b200_backward: model and stf
           H2D, Shot<id>.bin read + H2D, gradients D2H inside the timed region).
  roofline: dominant kernel (largest share of the step), algorithmic bytes per launch (DESIGN.md section 3:
           60 B/cell forward, 64 B/box-cell reverse+imaging, 60 B/cell adjoint, + CPML strips) / CUDA-event time per
           launch / measured HBM peak (MEASURED_PEAKS.json); traffic = ncu dram bytes per launch (profiles/traffic.json).
  cpu_baseline: the CPU oracle port (oracle/) on the host cores, bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SHOTS_PER_GPU = 30
NSTEPS = 2000
METRIC = "shot_gradients_per_s"
UNIT = "shot-gradients/s"
WORKLOAD = ("C2: 2-D elastic FWI gradient (fwi_op calc_id 1), 134x384 layered model padded to 224x448, "
            "379 receivers, 2000 steps, 30 shots per GPU")


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic():
    """dram bytes per launch of the dominant kernels from the committed ncu capture (profiles/traffic.json)."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    return json.load(open(p)) if os.path.exists(p) else {}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu=0):
        self.gpu, self.rows, self.proc = gpu, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.gpu)], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *a):
        if self.proc:
            time.sleep(0.15)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
                for n, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm)}


def make_case(n_gpus):
    from fwiflow.jl_b200 import synthetic
    if n_gpus == 1:
        return synthetic.case_c2(nshots=SHOTS_PER_GPU, nSteps=NSTEPS)
    return _multi_case(n_gpus)


def _multi_case(n_gpus):
    """30*N shots on the C2 grid: the weak-scaling survey (same receivers, sources spread over the line)."""
    from fwiflow.jl_b200 import synthetic
    from fwiflow.jl_b200.utils import sourceGene
    c = synthetic.case_c2(nshots=SHOTS_PER_GPU, nSteps=NSTEPS)
    n = SHOTS_PER_GPU * n_gpus
    c.x_src = np.round(np.linspace(4, c.nx - 5, n)).astype(np.int64)
    c.z_src = np.full(n, 2, dtype=np.int64)
    c.stf = np.repeat(sourceGene(4.5, NSTEPS, 0.0025), n, axis=0)
    return c


def cpu_baseline(threads=None, shots=4, steps=400):
    """CPU oracle port on the host cores: gradient (forward + backward) of `shots` C2 shots x `steps` steps."""
    from fwiflow.jl_b200 import synthetic
    from oracle import oracle_py as op
    cores = threads or os.cpu_count() or 1
    c = synthetic.case_c2(nshots=shots, nSteps=steps)
    para = c.write_files(tempfile.mkdtemp(prefix="bench_cpu_"))
    ids = np.arange(shots, dtype=np.int32)
    lam, mu, rho = c.moduli("true")
    lam0, mu0, rho0 = c.moduli("init")
    op.oracle_cufd(2, lam, mu, rho, c.stf, ids, para, threads=cores)
    t0 = time.perf_counter()
    op.oracle_cufd(1, lam0, mu0, rho0, c.stf, ids, para, threads=cores)
    dt = time.perf_counter() - t0
    cell_steps = shots * c.nz_pad * c.nx_pad * (steps - 1)       # per-cell time indices with fwd + bwd work
    rate = cell_steps / dt
    full = c.nz_pad * c.nx_pad * (NSTEPS - 1)                    # one full C2 shot-gradient
    return {"value": rate / full, "unit": UNIT, "cores": int(cores), "kind": "port",
            "sample": f"oracle/fwi_oracle.cpp (OpenMP), calc_id 1 on {shots} C2 shots x {steps} steps "
                      f"({dt:.1f} s), scaled by cell-updates to 2000-step shots",
            "cell_updates_per_s": 2 * rate}


def run_reference(args):
    """--impl reference: the reference's own implementation of the path.  It has NO CPU implementation (its
    TF kernels are DEVICE_CPU wrappers that call CUDA), so this times its CUDA op rebuilt for sm_100
    (oracle/_ref/libCUFD_ref.so) through its host-buffer entry point cufd(); if that library cannot run
    (no GPU / not built) the CPU oracle port is timed instead."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle_py as op
    c = make_case(1)
    ids = np.arange(SHOTS_PER_GPU, dtype=np.int32)
    lam, mu, rho = c.moduli("true")
    lam0, mu0, rho0 = c.moduli("init")
    use_ref = op.ref_available()
    if use_ref:
        try:
            import torch
            use_ref = torch.cuda.is_available()
        except Exception:
            use_ref = False
    line = {"metric": METRIC, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "data": "synthetic", "impl": "reference",
            "dtype": "f32", "config": {"workload": WORKLOAD, "l2": "working set 30 shots x 16 MB > L2"}}
    if use_ref:
        para = c.write_files(tempfile.mkdtemp(prefix="bench_ref_"))
        op.ref_cufd(2, lam, mu, rho, c.stf, ids, para)
        times = []
        with ClockSampler(0) as cs:
            for k in range(args.warmup + args.steps):
                t0 = time.perf_counter()
                op.ref_cufd(1, lam0, mu0, rho0, c.stf, ids, para)
                if k >= args.warmup:
                    times.append(time.perf_counter() - t0)
        dt = float(np.sum(times))
        v = SHOTS_PER_GPU * args.steps / dt
        line.update(value=v, ms_per_step=1e3 * dt / args.steps, clocks=cs.summary(),
                    cpu_baseline={"value": v, "unit": UNIT, "kind": "reference", "cores": 1,
                                  "sample": "unmodified reference cufd() (CUDA op rebuilt for sm_100, 1 GPU, 1 host thread), "
                                            "all 30 C2 shots x 2000 steps per step, host buffers + its own file I/O"},
                    e2e={"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                    note="the reference has no CPU implementation of this path; its CUDA op runs on 1 GPU regardless of --gpus")
    else:
        cb = cpu_baseline()
        cb["kind"] = "port"
        line.update(value=cb["value"], ms_per_step=1e3 / cb["value"] * SHOTS_PER_GPU, cpu_baseline=cb,
                    e2e={"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                    note="oracle/_ref not runnable here: CPU oracle port timed instead")
    print(json.dumps(line), flush=True)


def run_ours(args):
    import torch
    import torch.distributed as dist
    from fwiflow.jl_b200 import dist as fdist
    from fwiflow.jl_b200 import ops

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the FWI path has no CPU fallback")
    # stdout carries exactly ONE JSON line: native libraries that write to fd 1 (NCCL prints its version banner there)
    # are pointed at stderr for the duration of the run; the line itself goes to the saved descriptor
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n_gpus = world
    c = make_case(n_gpus)
    workdir = tempfile.mkdtemp(prefix=f"bench_r{rank}_")
    para = c.write_files(workdir)
    all_ids = np.arange(c.nShots, dtype=np.int32)
    my_ids = fdist.shard_shots(all_ids, rank, world)
    lam, mu, rho = c.moduli("true")
    lam0, mu0, rho0 = c.moduli("init")

    plan = ops.Plan(para, my_ids, gpu_id=local)
    plan.set_stf(c.stf)
    plan.set_model(lam, mu, rho)
    plan.run(2)                       # observed data of this rank's shots (synthetic, true model)
    plan.write_obs_files()
    plan.set_model(lam0, mu0, rho0)
    plan.load_obs_files()             # inputs now resident in HBM
    result = plan.result_tensor()
    stream = torch.cuda.current_stream()

    def step():
        plan.run(1, stream=stream.cuda_stream, sync=False)
        if world > 1:
            dist.all_reduce(result, op=dist.ReduceOp.SUM)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    l0 = plan.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as cs:
        barrier()
        e0.record(stream)
        for _ in range(args.steps):
            step()
        e1.record(stream)
        barrier()
    ms = e0.elapsed_time(e1)
    launches = plan.launch_count() - l0
    if world > 1:
        t = torch.tensor([ms], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        lt = torch.tensor([launches], device="cuda", dtype=torch.int64)
        dist.all_reduce(lt, op=dist.ReduceOp.SUM)
        launches = int(lt.item())
    total_shots = c.nShots * args.steps
    value = total_shots / (ms * 1e-3)
    cells = c.nz_pad * c.nx_pad
    cell_updates = 2.0 * total_shots * cells * (NSTEPS - 1) / (ms * 1e-3)   # forward + backward updates

    # ---- kernel roofline (rank 0): CUDA events around each hot kernel, on this stream ----
    peak, peak_src = measured_peak()
    traffic = ncu_traffic()
    kernels = []
    names = {1: "fwd_step_kernel<save_frames>", 2: "rev_image_kernel", 3: "adj_step_kernel"}
    per_cell = {1: "60 + 32 (fz + fx) B per cell + frame quads", 2: "64 B per inner-box cell",
                3: "60 + 64 (fz + fx) B per cell"}
    per_step_launch = {1: NSTEPS - 1, 2: NSTEPS - 1, 3: NSTEPS}
    nb = max(1, -(-len(my_ids) // plan.batch))
    if rank == 0:
        for which in (1, 2, 3):
            k_ms, k_bytes = plan.time_kernel(which, iters=200, stream=stream.cuda_stream)
            ach = k_bytes / (k_ms * 1e-3) / 1e9
            kernels.append({"kernel": names[which], "ms_per_launch": k_ms, "alg_bytes_per_launch": k_bytes,
                            "alg_bytes_rule": per_cell[which],
                            "achieved_gbs": ach, "frac": ach / peak, "launches_per_step": per_step_launch[which] * nb,
                            "share_of_step": per_step_launch[which] * nb * k_ms / (ms / args.steps),
                            "traffic": traffic.get(names[which])})
        dom = max(kernels, key=lambda k: k["share_of_step"])
        roofline = {"bound": "hbm", "kernel": dom["kernel"], "achieved": dom["achieved_gbs"], "peak": peak,
                    "unit": "GB/s", "frac": dom["frac"], "traffic": dom["traffic"], "peak_source": peak_src,
                    "timing": "CUDA events on the launching stream, 200 back-to-back launches, batch of "
                              f"{min(plan.batch, len(my_ids))} shots per launch"}
    barrier()

    # ---- end to end through the reference-facing host-buffer entry point ----
    ops.fwi_op_grad(lam0, mu0, rho0, c.stf, local, my_ids, para)         # warm the plan cache
    barrier()
    t0 = time.perf_counter()
    e2e_steps = max(1, min(args.steps, 3))
    for _ in range(e2e_steps):
        gl, gm, gd, gs = ops.fwi_op_grad(lam0, mu0, rho0, c.stf, local, my_ids, para)
        if world > 1:
            buf = torch.from_numpy(np.stack([gl, gm, gd])).cuda()
            dist.all_reduce(buf, op=dist.ReduceOp.SUM)
            buf.cpu()
    barrier()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_s], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    nrec = c.nrec
    h2d = 3 * cells * 8 + len(my_ids) * NSTEPS * 4 + len(my_ids) * nrec * NSTEPS * 4
    d2h = (3 * cells + 1) * 4 + len(my_ids) * NSTEPS * 4
    e2e = {"value": c.nShots * e2e_steps / e2e_s, "unit": UNIT, "h2d_bytes_per_step": int(h2d),
           "d2h_bytes_per_step": int(d2h), "steps": e2e_steps,
           "call": "fwi_b200_backward(host buffers): model+stf H2D, Data/Shot<id>.bin read + H2D, run, gradients D2H"}

    if rank == 0:
        cb = cpu_baseline() if (n_gpus == 1 and not args.no_cpu) else None
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": n_gpus, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": WORKLOAD, "shots_total": int(c.nShots), "batch": int(plan.batch),
                           "l2": "inputs larger than L2: 30 shots x 36 planes x 0.41 MB = 443 MB of wavefield state per GPU",
                           "parallelism": f"shots sharded over {n_gpus} GPU(s), one all-reduce per gradient"},
                "cell_updates_per_s": cell_updates, "gpu_launches": int(launches), "e2e": e2e, "roofline": roofline,
                "kernels": kernels, "clocks": cs.summary()}
        if cb:
            line["cpu_baseline"] = cb
        sys.stdout.flush()
        os.write(json_fd, (json.dumps(line) + "\n").encode())
    plan.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
