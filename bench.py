#!/usr/bin/env python
"""bench.py -- shot-gradient throughput of the FWI hot path on B200.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's own op (oracle/_ref), driven by rank 0

Headline workload (BASELINE.json configs[1], "C2"): Marmousi-sized 134x384 model (224x448 padded), 379 receivers,
2000 time steps, 30 shots PER GPU (weak scaling), forward + adjoint gradient as `fwi_op`'s gradient kernel computes
it (calc_id 1).  One "step" = one gradient evaluation of the rank's 30 shots.
  value   : shot-gradients/s of the whole job, inputs (model, stf, observed data) resident in HBM, timed with CUDA
            events on the launching stream, barrier + synchronize on both sides, max over ranks; for N > 1 every step
            ends with the NCCL all-reduce of [grad_lambda|grad_mu|grad_den|misfit].
  e2e     : same metric through the reference-facing C-ABI call with HOST buffers (fwi_b200_backward: model and stf
            H2D, Shot<id>.bin read + H2D, gradients D2H inside the timed region).
  roofline: dominant kernel (largest share of the step), algorithmic bytes per launch /
            CUDA-event time per launch / measured HBM peak (MEASURED_PEAKS.json); `traffic` = ncu dram bytes per launch
            from the committed capture (profiles/traffic.json), `dram_frac` = that traffic / time / peak.
  configs : the DRAM-bound BASELINE config C3 (1088x3064 padded, 25 shots per GPU = 200 shots on 8 GPUs, 4000 steps)
            measured in the same run: gradient time, per-kernel times and fractions, whole-gradient fraction.
  strong  : strong scaling -- a FIXED survey split over the N ranks: (i) configs[1] literally, 30 C2 shots in total;
            (ii) C3, 200 shots (record shortened to 1000 steps so that the N = 1 point fits the run).
  multi_c_abi (N > 1): rank 0 alone drives all N GPUs through ONE C-ABI call (fwi_b200_gradient_multi: one host
            thread per device, ncclAllReduce inside the library) with host buffers, while the other ranks idle.
  cpu_baseline: the CPU oracle port (oracle/) on the host cores, bounded sample of the same workload (N = 1 only).
The reference arm runs the reference's own CUDA op the reference's own way: one cufd() per gpu_id, concurrently (one
host process per GPU), contiguous shot ranges (test/TestFWI.jl:58-69), and one C3 shot for the `configs` record.
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
C3_SHOTS_PER_GPU = 25          # 200 shots on 8 GPUs (BASELINE.json configs[2])
C3_NSTEPS = 4000
C3_WORKLOAD = ("C3: 1000x3000 layered model padded to 1088x3064, 2994 receivers, 4000 steps, 25 shots per GPU "
               "(the per-GPU share of 200 shots on 8 GPUs), gradient with boundary-frame checkpoints")
STRONG_C3_SHOTS = 200
STRONG_C3_NSTEPS = 1000


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic():
    """dram bytes per launch of the hot kernels from the committed ncu captures (profiles/traffic.json):
    {"c2": {kernel: bytes at 30 shots}, "c3": {kernel: bytes at 8 shots}}."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if not os.path.exists(p):
        return {}
    try:
        raw = json.load(open(p))
    except (OSError, ValueError):
        return {}
    # per config: kernel name -> bytes; notes, shot counts and anything else that is not a number is left out
    return {cfg: {k: float(v) for k, v in d.items() if isinstance(v, (int, float)) and not k.startswith("_")}
            for cfg, d in raw.items() if isinstance(d, dict)}


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


# ---- workloads -------------------------------------------------------------------------------------------------------
def c2_case(nshots):
    """nshots sources on the C2 grid: the 30-shot survey of configs[1] for nshots == 30, else the same receivers with
    the sources spread over the line (weak scaling: 30 per GPU)."""
    from fwiflow.jl_b200 import synthetic
    from fwiflow.jl_b200.utils import sourceGene
    c = synthetic.case_c2(nshots=SHOTS_PER_GPU, nSteps=NSTEPS)
    if nshots != SHOTS_PER_GPU:
        c.x_src = np.round(np.linspace(4, c.nx - 5, nshots)).astype(np.int64)
        c.z_src = np.full(nshots, 2, dtype=np.int64)
        c.stf = np.repeat(sourceGene(4.5, NSTEPS, 0.0025), nshots, axis=0)
    return c


def make_case(n_gpus):
    return c2_case(SHOTS_PER_GPU * n_gpus)


def c3_case(nshots, nsteps):
    from fwiflow.jl_b200 import synthetic
    return synthetic.case_c3(nshots=nshots, nSteps=nsteps)


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


# ---- the reference arm -----------------------------------------------------------------------------------------------
_REF_WORKER = r"""
import json, os, sys, time
sys.path.insert(0, sys.argv[1])
import numpy as np
from oracle import oracle_py as op
job = np.load(sys.argv[2], allow_pickle=False)
ids = job["ids"].astype(np.int32); gpu = int(job["gpu"]); para = str(job["para"]); reps = int(job["reps"])
op.ref_cufd(2, job["lam"], job["mu"], job["rho"], job["stf"], ids, para, gpu_id=gpu)      # its own Shot<id>.bin
print("READY", flush=True)
sys.stdin.readline()                                                                    # start gun from the parent
times = []
for k in range(reps):
    t0 = time.perf_counter()
    op.ref_cufd(1, job["lam0"], job["mu0"], job["rho0"], job["stf"], ids, para, gpu_id=gpu)
    times.append(time.perf_counter() - t0)
    print("T %.6f %.6f" % (t0, times[-1]), flush=True)
"""


def reference_ranges(nshots, ngpu):
    """The reference's own split (test/TestFWI.jl:58-69): shot_id_points = trunc(LinRange(1, nShots, nGpus + 1)),
    GPU i gets the INCLUSIVE range points[i]..points[i+1] -- boundary shots are processed by two GPUs (1-based)."""
    pts = np.trunc(np.linspace(1, nshots, ngpu + 1)).astype(int)
    return [np.arange(pts[i], pts[i + 1] + 1) - 1 for i in range(ngpu)]      # 0-based ids


def run_reference(args):
    """--impl reference: the reference's own implementation of the path.  It has NO CPU implementation (its TF
    kernels are DEVICE_CPU wrappers that call CUDA), so this times its CUDA op rebuilt for sm_100
    (oracle/_ref/libCUFD_ref.so) through its host-buffer entry point cufd() -- on N GPUs the way the reference's own
    driver does it: one cufd() per gpu_id at the same time (one host process per GPU stands in for TensorFlow's
    inter-op threads), contiguous inclusive shot ranges.  If that library cannot run (no GPU / not built) the CPU
    oracle port is timed instead."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle_py as op
    n = max(1, args.gpus)
    c = make_case(n)
    lam, mu, rho = c.moduli("true")
    lam0, mu0, rho0 = c.moduli("init")
    use_ref = op.ref_available()
    if use_ref:
        try:
            import torch
            use_ref = torch.cuda.is_available() and torch.cuda.device_count() >= n
        except Exception:
            use_ref = False
    line = {"metric": METRIC, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "data": "synthetic", "impl": "reference",
            "dtype": "f32", "config": {"workload": WORKLOAD, "l2": "working set 30 shots x 16 MB > L2"}}
    if not use_ref:
        cb = cpu_baseline()
        cb["kind"] = "port"
        line.update(value=cb["value"], ms_per_step=1e3 / cb["value"] * SHOTS_PER_GPU, cpu_baseline=cb,
                    e2e={"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                    note="oracle/_ref not runnable here: CPU oracle port timed instead")
        print(json.dumps(line), flush=True)
        return
    wd = tempfile.mkdtemp(prefix="bench_ref_")
    para = c.write_files(wd)
    ranges = reference_ranges(c.nShots, n) if n > 1 else [np.arange(c.nShots)]
    reps = args.warmup + args.steps
    procs = []
    for g, ids in enumerate(ranges):
        job = os.path.join(wd, f"job{g}.npz")
        np.savez(job, ids=ids, gpu=g, para=para, reps=reps, lam=lam, mu=mu, rho=rho, lam0=lam0, mu0=mu0, rho0=rho0,
                 stf=c.stf)
        procs.append(subprocess.Popen([sys.executable, "-c", _REF_WORKER, ROOT, job], stdin=subprocess.PIPE,
                                      stdout=subprocess.PIPE, text=True))
    for p in procs:
        assert p.stdout.readline().strip() == "READY", "reference worker failed"
    with ClockSampler(0) as cs:
        for p in procs:
            p.stdin.write("go\n"); p.stdin.flush()
        spans = []
        for p in procs:
            rows = [ln.split() for ln in p.stdout if ln.startswith("T ")]
            p.wait()
            spans.append([(float(r[1]), float(r[2])) for r in rows])
    # one "step" = one gradient of the whole survey: from the earliest start to the latest end of repetition k
    step_s = [max(s[k][0] + s[k][1] for s in spans) - min(s[k][0] for s in spans) for k in range(args.warmup, reps)]
    dt = float(np.sum(step_s))
    v = c.nShots * args.steps / dt
    dup = int(sum(len(r) for r in ranges) - c.nShots)
    line.update(value=v, ms_per_step=1e3 * dt / args.steps, clocks=cs.summary(),
                cpu_baseline={"value": v, "unit": UNIT, "kind": "reference", "cores": n,
                              "sample": f"unmodified reference cufd() (CUDA op rebuilt for sm_100), {n} GPU(s), one host "
                                        f"process per GPU, all {c.nShots} C2 shots x 2000 steps per step, host buffers + its "
                                        f"own file I/O" + (f"; its inclusive ranges process {dup} boundary shot(s) twice "
                                                           "(test/TestFWI.jl:65), counted once" if dup else "")},
                e2e={"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                note="the reference has no CPU implementation of this path; this is its CUDA op, sharded its own way")
    # ---- the same op at the DRAM-bound config: ONE C3 shot x 4000 steps (its shots are sequential: per-shot time) ----
    try:
        c3 = c3_case(1, C3_NSTEPS)
        para3 = c3.write_files(tempfile.mkdtemp(prefix="bench_ref_c3_"))
        ids = np.array([0], np.int32)
        l3, m3, r3 = c3.moduli("true")
        op.ref_cufd(2, l3, m3, r3, c3.stf, ids, para3)
        l30, m30, r30 = c3.moduli("init")
        t0 = time.perf_counter()
        op.ref_cufd(1, l30, m30, r30, c3.stf, ids, para3)
        t3 = time.perf_counter() - t0
        cells = c3.nz_pad * c3.nx_pad
        line["configs"] = {"c3": {"workload": C3_WORKLOAD, "sample": "1 shot x 4000 steps through cufd() on 1 GPU "
                                  "(the reference runs the shots of a GPU one after the other)",
                                  "value": 1.0 / t3, "unit": UNIT + " per GPU", "s_per_shot_gradient": t3,
                                  "cell_updates_per_s": 2.0 * cells * (C3_NSTEPS - 1) / t3}}
    except Exception as e:  # the headline stays valid without it
        line["configs"] = {"c3": {"error": str(e)[:200]}}
    print(json.dumps(line), flush=True)


# ---- this repo's arm ---------------------------------------------------------------------------------------------------
class Dist:
    def __init__(self):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(self.local)
        self.cpu_group = None
        if self.world > 1:
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local))
            self.cpu_group = dist.new_group(backend="gloo")   # host-side barrier that leaves the GPUs idle

    def host_barrier(self):
        """Wait on the HOST only: a NCCL barrier is a kernel that spins on the waiting ranks' GPUs."""
        self.torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier(group=self.cpu_group)

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max(self, x):
        if self.world == 1:
            return x
        t = self.torch.tensor([x], device="cuda", dtype=self.torch.float64)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def sum_int(self, x):
        if self.world == 1:
            return int(x)
        t = self.torch.tensor([x], device="cuda", dtype=self.torch.int64)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return int(t.item())


def resident_plan(D, c, my_ids, workdir):
    """Plan of this rank's shots with model, stf and observed data (synthetic, true model) resident in HBM."""
    from fwiflow.jl_b200 import ops
    para = c.write_files(workdir)
    lam, mu, rho = c.moduli("true")
    plan = ops.Plan(para, my_ids, gpu_id=D.local)
    plan.set_stf(c.stf)
    plan.set_model(lam, mu, rho)
    plan.run(2)
    plan.write_obs_files()
    plan.set_model(*c.moduli("init"))
    plan.load_obs_files()
    return plan, para


def timed_gradients(D, plan, steps, warmup, sampler_gpu=None):
    """`steps` gradient evaluations (+ all-reduce for N > 1) between CUDA events on the launching stream, barrier +
    synchronize on both sides; returns (ms total = max over ranks, launches summed over ranks, clocks)."""
    torch, dist = D.torch, D.dist
    result = plan.result_tensor()
    stream = torch.cuda.current_stream()   # the created stream run_ours() made current: kernels, all-reduce, events
    assert stream.cuda_stream != 0

    def step():
        plan.run(1, stream=stream.cuda_stream, sync=False)
        if D.world > 1:
            dist.all_reduce(result, op=dist.ReduceOp.SUM)

    for _ in range(warmup):
        step()
    D.barrier()
    l0 = plan.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(D.local if sampler_gpu is None else sampler_gpu) as cs:
        D.barrier()
        e0.record(stream)
        for _ in range(steps):
            step()
        e1.record(stream)
        D.barrier()
    ms = D.max(e0.elapsed_time(e1))
    launches = D.sum_int(plan.launch_count() - l0)
    return ms, launches, cs.summary()


def kernel_table(plan, stream, nsteps, nbatches, step_ms, peak, traffic, iters):
    """CUDA-event time per launch of the step kernels of a gradient (back-to-back launches on `stream`), their
    algorithmic bytes, and the ncu dram bytes of the committed capture."""
    spec = [  # (C-ABI selector, name, launches per gradient and batch, rule)
        (1, "fwd_step_kernel<save_frames>", nsteps - 1, "60 + 32 (fz + fx) B per cell + frame quads"),
        (2, "rev_image_kernel", nsteps - 1, "64 B per inner-box cell"),
        (3, "adj_step_kernel", nsteps, "60 + 64 (fz + fx) B per cell"),
    ]
    rows = []
    for which, name, n_launch, rule in spec:
        k_ms, k_bytes = plan.time_kernel(which, iters=iters, stream=stream.cuda_stream)
        ach = k_bytes / (k_ms * 1e-3) / 1e9
        tr = traffic.get(name)
        rows.append({"kernel": name, "ms_per_launch": k_ms, "alg_bytes_per_launch": k_bytes, "alg_bytes_rule": rule,
                     "achieved_gbs": ach, "frac": ach / peak, "launches_per_step": n_launch * nbatches,
                     "share_of_step": n_launch * nbatches * k_ms / step_ms,
                     "traffic": tr, "dram_frac": (tr / (k_ms * 1e-3) / 1e9 / peak) if tr else None})
    return rows


def whole_gradient_alg_bytes(c, nshots, nsteps):
    """DESIGN.md section 3 / SURVEY.md 8d, every launch of a gradient: forward 60 + 32 (fz + fx), adjoint
    60 + 64 (fz + fx) per padded cell, reverse + imaging 64 per inner-box cell, per time index."""
    cells = c.nz_pad * c.nx_pad
    box = (c.nz_pad - c.nPad - 2 * c.nPml) * (c.nx_pad - 2 * c.nPml)
    fz, fx = 2.0 * c.nPml / c.nz_pad, 2.0 * c.nPml / c.nx_pad
    per_index = cells * (60 + 32 * (fz + fx)) + cells * (60 + 64 * (fz + fx)) + box * 64.0
    return float(nshots) * per_index * (nsteps - 1)


def timelapse_record(c, gpus, reps):
    """BASELINE configs[3] (C4): 6 surveys (baseline + 5 monitors with a growing -5 % lambda / -1 % rho anomaly, each with
    its own para file and Data directory, 30 shots x 2000 steps on the C2 grid) evaluated at the baseline model through
    fwi_b200_timelapse with host buffers: survey i on gpus[i % len(gpus)], cached plans stay resident between calls."""
    from fwiflow.jl_b200 import ops
    lam, mu, rho = c.moduli("true")
    z, x = np.mgrid[0:c.nz_pad, 0:c.nx_pad]
    ids = np.arange(SHOTS_PER_GPU, dtype=np.int32)
    stf = c.stf[:SHOTS_PER_GPU]
    surveys = []
    for k in range(6):
        r = 6.0 + 3.0 * k
        blob = np.exp(-(((z - (c.nPml + 0.55 * c.nz)) / r) ** 2 + ((x - (c.nPml + 0.5 * c.nx)) / (2.0 * r)) ** 2)) if k else 0.0
        lam_k, rho_k = lam * (1.0 - 0.05 * blob), rho * (1.0 - 0.01 * blob)
        ck = c2_case(SHOTS_PER_GPU)
        para = ck.write_files(tempfile.mkdtemp(prefix=f"bench_c4_s{k}_"))
        ops.fwi_obs_op(lam_k, mu, rho_k, stf, gpus[k % len(gpus)], ids, para)        # this survey's observations
        surveys.append((para, lam, mu, rho))                                          # every survey evaluated at the baseline
    out = ops.timelapse(surveys, stf, gpus, ids)                                     # plans + observations warm
    t0 = time.perf_counter()
    for _ in range(reps):
        out = ops.timelapse(surveys, stf, gpus, ids)
    dt = (time.perf_counter() - t0) / reps
    js = [o[0] for o in out]
    return {"workload": "C4: time-lapse, 6 surveys x 30 shots x 2000 steps on the C2 grid, one fwi_b200_timelapse call "
                        f"(host buffers) on {len(gpus)} GPU(s)", "surveys": 6, "shots_per_survey": SHOTS_PER_GPU,
            "value": 6 * SHOTS_PER_GPU / dt, "unit": UNIT, "surveys_per_s": 6 / dt, "s_per_call": dt, "steps": reps,
            "misfits": js, "monitor_misfits_grow": bool(js[0] == 0.0 and all(js[k + 1] > js[k] for k in range(5)))}


def run_ours(args):
    import torch
    from fwiflow.jl_b200 import dist as fdist
    from fwiflow.jl_b200 import ops

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the FWI path has no CPU fallback")
    # stdout carries exactly ONE JSON line: native libraries that write to fd 1 (NCCL prints its version banner there)
    # are pointed at stderr for the duration of the run; the line itself goes to the saved descriptor
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    D = Dist()
    rank, world, local = D.rank, D.world, D.local
    n_gpus = world
    warmup = max(args.warmup, 3)
    peak, peak_src = measured_peak()
    traffic = ncu_traffic()
    # ONE created stream carries everything that is timed: the plan's kernels, the NCCL all-reduce and the CUDA events.
    # (torch's default stream is CUDA's legacy stream, handle 0 -- the C ABI reads a null handle as "the plan's own
    # stream", and work there is not ordered with events or collectives on the default stream.)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)

    # ================= headline: C2, 30 shots per GPU (weak) =================
    c = make_case(n_gpus)
    all_ids = np.arange(c.nShots, dtype=np.int32)
    my_ids = fdist.shard_shots(all_ids, rank, world)
    plan, para = resident_plan(D, c, my_ids, tempfile.mkdtemp(prefix=f"bench_r{rank}_"))
    ms, launches, clocks = timed_gradients(D, plan, args.steps, warmup)
    total_shots = c.nShots * args.steps
    value = total_shots / (ms * 1e-3)
    cells = c.nz_pad * c.nx_pad
    cell_updates = 2.0 * total_shots * cells * (NSTEPS - 1) / (ms * 1e-3)   # forward + backward updates
    batch = int(plan.batch)
    nb = max(1, -(-len(my_ids) // plan.batch))
    kernels, roofline = [], None
    if rank == 0:
        kernels = kernel_table(plan, stream, NSTEPS, nb, ms / args.steps, peak, traffic.get("c2", {}), 200)
        dom = max(kernels, key=lambda k: k["share_of_step"])
        roofline = {"bound": "hbm", "kernel": dom["kernel"], "achieved": dom["achieved_gbs"], "peak": peak,
                    "unit": "GB/s", "frac": dom["frac"], "traffic": dom["traffic"], "dram_frac": dom["dram_frac"],
                    "peak_source": peak_src,
                    "traffic_source": "ncu --set full capture of the same kernels (profiles/traffic.json), not re-measured here",
                    "whole_gradient_frac": whole_gradient_alg_bytes(c, len(my_ids), NSTEPS) * world /
                                           (ms / args.steps * 1e-3) / 1e9 / peak / world,
                    "timing": "CUDA events on the launching stream, 200 back-to-back launches, batch of "
                              f"{min(plan.batch, len(my_ids))} shots per launch"}
    D.barrier()

    # ---- end to end through the reference-facing host-buffer entry point (per rank, + all-reduce for N > 1) ----
    lam0, mu0, rho0 = c.moduli("init")
    ops.fwi_op_grad(lam0, mu0, rho0, c.stf, local, my_ids, para)         # warm the plan cache
    D.barrier()
    t0 = time.perf_counter()
    e2e_steps = max(1, min(args.steps, 3))
    for _ in range(e2e_steps):
        gl, gm, gd, gs = ops.fwi_op_grad(lam0, mu0, rho0, c.stf, local, my_ids, para)
        if world > 1:
            buf = torch.from_numpy(np.stack([gl, gm, gd])).cuda()
            D.dist.all_reduce(buf, op=D.dist.ReduceOp.SUM)
            buf.cpu()
    D.barrier()
    e2e_s = D.max(time.perf_counter() - t0)
    nrec = c.nrec
    h2d = 3 * cells * 8 + len(my_ids) * NSTEPS * 4 + len(my_ids) * nrec * NSTEPS * 4
    d2h = (3 * cells + 1) * 4 + len(my_ids) * NSTEPS * 4
    e2e = {"value": c.nShots * e2e_steps / e2e_s, "unit": UNIT, "h2d_bytes_per_step": int(h2d),
           "d2h_bytes_per_step": int(d2h), "steps": e2e_steps,
           "call": "fwi_b200_backward(host buffers): model+stf H2D, Data/Shot<id>.bin read + H2D, run, gradients D2H"}
    ops.release()
    plan.close()
    D.barrier()

    # ---- N > 1: the whole node through ONE C-ABI call from rank 0 (NCCL inside the library) ----
    multi = None
    if world > 1 and not args.no_multi:
        D.host_barrier()          # every rank idle, no collective kernel resident on any GPU
        if rank == 0:
            try:
                wd = tempfile.mkdtemp(prefix="bench_multi_")
                para_m = c.write_files(wd)
                lam, mu, rho = c.moduli("true")
                gpus = list(range(world))
                for g in gpus:   # observations of every shot, written by the devices that will read them
                    ops.fwi_obs_op(lam, mu, rho, c.stf, g, all_ids[g::world], para_m)
                ops.fwi_op_and_grad_multi(lam0, mu0, rho0, c.stf, gpus, all_ids, para_m)     # plans, communicators
                t0 = time.perf_counter()
                for _ in range(e2e_steps):
                    ops.fwi_op_and_grad_multi(lam0, mu0, rho0, c.stf, gpus, all_ids, para_m)
                dt_m = time.perf_counter() - t0
                multi = {"value": c.nShots * e2e_steps / dt_m, "unit": UNIT, "steps": e2e_steps,
                         "call": f"fwi_b200_gradient_multi(host buffers) from ONE process on {world} GPUs: one host thread "
                                 "per device, ncclAllReduce of the packed gradients inside the library, single D2H"}
                ops.release()
            except Exception as e:
                multi = {"error": str(e)[:300]}
        D.host_barrier()

    # ================= C4: time-lapse, baseline + 5 monitor surveys in ONE C-ABI call (rank 0 drives every GPU) =================
    c4 = None
    if not args.no_c4:
        D.host_barrier()
        if rank == 0:
            try:
                c4 = timelapse_record(c, list(range(world)), e2e_steps)
            except Exception as e:
                c4 = {"error": str(e)[:300]}
            ops.release()
        D.host_barrier()

    # ================= strong scaling (i): configs[1] literally -- 30 C2 shots over N ranks =================
    strong = {}
    if world > 1:
        cs = c2_case(SHOTS_PER_GPU)
        ids_s = fdist.shard_shots(np.arange(cs.nShots, dtype=np.int32), rank, world)
        p_s, para_s = resident_plan(D, cs, ids_s, tempfile.mkdtemp(prefix=f"bench_s2_r{rank}_"))
        ms_s, _, _ = timed_gradients(D, p_s, args.steps, 3)
        gi = ops.grid_info(para_s)
        ntl = gi["tiles_z"] * gi["tiles_x"]
        strong["c2_30_shots"] = {"workload": "configs[1]: 30 C2 shots x 2000 steps in total, split over the ranks",
                                 "shots_total": 30, "shots_per_rank": [int(len(fdist.shard_shots(np.arange(30), r, world))) for r in range(world)],
                                 "value": 30 * args.steps / (ms_s * 1e-3), "unit": UNIT, "ms_per_step": ms_s / args.steps,
                                 "work_items_per_launch_rank0": int(len(ids_s)) * ntl if rank == 0 else None}
        p_s.close()
        D.barrier()
    else:
        strong["c2_30_shots"] = {"workload": "configs[1]: 30 C2 shots x 2000 steps in total, split over the ranks",
                                 "shots_total": 30, "shots_per_rank": [30], "value": value, "unit": UNIT,
                                 "ms_per_step": ms / args.steps, "note": "N = 1: identical to the headline run"}

    # ================= C3: the DRAM-bound config, 25 shots per GPU x 4000 steps =================
    configs = {}
    if not args.no_c3:
        c3 = c3_case(C3_SHOTS_PER_GPU * world, C3_NSTEPS)
        ids3 = fdist.shard_shots(np.arange(c3.nShots, dtype=np.int32), rank, world)
        p3, _ = resident_plan(D, c3, ids3, tempfile.mkdtemp(prefix=f"bench_c3_r{rank}_"))
        ms3, launches3, clocks3 = timed_gradients(D, p3, 1, 1)
        cells3 = c3.nz_pad * c3.nx_pad
        rec = {"workload": C3_WORKLOAD, "shots_total": int(c3.nShots), "batch": int(p3.batch), "steps": 1, "warmup": 1,
               "value": c3.nShots / (ms3 * 1e-3), "unit": UNIT, "ms_per_gradient": ms3,
               "cell_updates_per_s": 2.0 * c3.nShots * cells3 * (C3_NSTEPS - 1) / (ms3 * 1e-3),
               "gpu_launches": launches3, "clocks": clocks3}
        if rank == 0:
            nb3 = max(1, -(-len(ids3) // p3.batch))
            # the committed ncu capture is of an 8-shot launch: scale its dram bytes to this batch
            t3 = {k: v * min(p3.batch, len(ids3)) / 8.0 for k, v in traffic.get("c3", {}).items()}
            rec["kernels"] = kernel_table(p3, stream, C3_NSTEPS, nb3, ms3, peak, t3, 40)
            rec["whole_gradient_frac"] = whole_gradient_alg_bytes(c3, len(ids3), C3_NSTEPS) / (ms3 * 1e-3) / 1e9 / peak
            rec["whole_gradient_rule"] = ("every launch of the gradient (three step kernels, residual, finalize, memsets) "
                                          "against 60 + 32 (fz+fx) forward + 60 + 64 (fz+fx) adjoint per cell + 64 per "
                                          "box cell reverse/imaging, per time index")
        configs["c3"] = rec
        p3.close()
        D.barrier()

        # ---- strong scaling (ii): C3, 200 shots split over the ranks (record shortened to 1000 steps) ----
        c3s = c3_case(STRONG_C3_SHOTS, STRONG_C3_NSTEPS)
        ids3s = fdist.shard_shots(np.arange(c3s.nShots, dtype=np.int32), rank, world)
        p3s, _ = resident_plan(D, c3s, ids3s, tempfile.mkdtemp(prefix=f"bench_c3s_r{rank}_"))
        ms3s, _, _ = timed_gradients(D, p3s, 1, 1)
        strong["c3_200_shots"] = {"workload": f"configs[2]: C3 grid, {STRONG_C3_SHOTS} shots in total split over the ranks, "
                                              f"{STRONG_C3_NSTEPS} of the 4000 steps (same per-step work; keeps the N = 1 point short)",
                                  "shots_total": STRONG_C3_SHOTS, "shots_per_rank": int(len(ids3s)), "batch": int(p3s.batch),
                                  "value": STRONG_C3_SHOTS / (ms3s * 1e-3), "unit": UNIT + f" ({STRONG_C3_NSTEPS}-step shots)",
                                  "ms_per_gradient": ms3s,
                                  "cell_updates_per_s": 2.0 * STRONG_C3_SHOTS * cells3 * (STRONG_C3_NSTEPS - 1) / (ms3s * 1e-3),
                                  "steps": 1, "warmup": 1}
        p3s.close()
        D.barrier()

    if rank == 0:
        cb = cpu_baseline() if (n_gpus == 1 and not args.no_cpu) else None
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": n_gpus, "steps": args.steps,
                "warmup": warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": WORKLOAD, "shots_total": int(c.nShots), "batch": batch,
                           "l2": "inputs larger than L2: 30 shots x 36 planes x 0.41 MB = 443 MB of wavefield state per GPU",
                           "parallelism": f"shots sharded over {n_gpus} GPU(s), one all-reduce per gradient"},
                "cell_updates_per_s": cell_updates, "gpu_launches": int(launches), "e2e": e2e, "roofline": roofline,
                "kernels": kernels, "clocks": clocks, "configs": configs, "strong": strong}
        if c4:
            line["configs"]["c4"] = c4
        if multi:
            line["multi_c_abi"] = multi
        if cb:
            line["cpu_baseline"] = cb
        sys.stdout.flush()
        os.write(json_fd, (json.dumps(line) + "\n").encode())
    if world > 1:
        D.dist.barrier()
        D.dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-c3", action="store_true", help="skip the C3 config record and the C3 strong-scaling record")
    ap.add_argument("--no-multi", action="store_true", help="skip the single-process multi-GPU C-ABI leg (N > 1)")
    ap.add_argument("--no-c4", action="store_true", help="skip the time-lapse (C4) record")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
