#!/usr/bin/env python
"""bench.py -- CPD EM-iterations/sec at N = M = 100k 3-D points on N B200s (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--points 100000]
  N > 1 is launched by torchrun (one rank per GPU); rank 0 prints ONE JSON line.

A "step" is one EM iteration (probreg/cpd.py:111-113: transform -> E-step -> M-step) of rigid CPD
on the synthetic workload of BASELINE.md section 3 (anisotropic box, 30 degree rotation, 0.01 noise),
w = 0, sigma2 auto-initialised, update_scale=True.  Strong scaling: the cloud is fixed and the
targets are sharded over the ranks (one 32-double NCCL all-reduce per iteration).

  value   : iterations/s with both clouds resident in HBM; K iterations timed back to back with a
            CUDA-event pair per iteration on the library's stream, L2 flushed (256 MiB memset)
            between iterations outside the event pairs; max over ranks.
  e2e     : the same metric through the public API with HOST buffers: each step is one
            RigidCPD.registration(target, maxiter=1) -- H2D of both clouds from pinned memory, sigma2
            init, one EM iteration, D2H of the MstepResult -- timed with the host clock around it.
  roofline: the fused E-step (pass 1 + pass 2 kernels) against the FP32 issue roofline it is bound
            by, plus the HBM view (the kernel moves ~10 MB per iteration: P is never materialised).
  cpu_baseline / --impl reference: the oracle port of the reference's numpy path (oracle/cpd_oracle.py)
            on the host cores, on a bounded column sample of the same workload, extrapolated.
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
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "cpd_em_iterations_per_sec"
UNIT = "it/s"
FLOP_PER_PAIR_ITER = 29.0     # pass 1 = 3D+2 = 11, pass 2 = 5D+3 = 18 at D = 3


def hbm_bytes_per_iter(n_local, m):
    # algorithmic traffic of the fused E-step, per iteration and rank:
    # pass 1 reads 16 B/target + 16 B/source, writes 8 B/target; pass 2 reads 32 B/target + 16 B/source
    # and writes 32 B/source (p1, px in FP64).
    return 16 * n_local + 16 * m + 8 * n_local + 32 * n_local + 16 * m + 32 * m


class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons (B200_PROFILING.md recipe).  ONE poller (rank 0) for all the
    job's GPUs, started before the warm-up so that NVML start-up (slow with 8 GPUs, and it takes driver
    locks) does not land inside the timed region; samples taken while the GPUs are busy are kept."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,utilization.gpu,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, n_gpus):
        self.n_gpus = n_gpus
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), line.strip()))

    def count(self):
        return len(self.rows)

    def stop(self, t0=None, t1=None):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons, power, timed = [], [], set(), [], 0
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ts, r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 9:
                continue
            try:
                idx = int(f[0])
                try:
                    util = float(f[4])
                except ValueError:
                    util = 100.0                          # utilisation not reported: keep the sample
                if idx >= self.n_gpus or util < 50.0:      # keep samples taken under load on this job's GPUs
                    continue
                sm.append(float(f[1])); smax.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            if t0 is not None and t0 <= ts <= t1 + 0.11:
                timed += 1
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "power_w_max": max(power) if power else None, "samples_under_load": len(sm),
                "samples_in_timed_region": timed, "reasons": sorted(reasons)}


def workload(points):
    from probreg_b200.synthetic import synthetic_pair
    return synthetic_pair(points, "rigid")


# ---------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference's numpy path on a bounded sample
# ---------------------------------------------------------------------------------------------
def cpu_sample(points, cols, repeats=1):
    """One EM iteration of the reference algorithm, E-step on `cols` of the `points` target columns
    (exact per column, cpd.py:80-87), extrapolated to all columns; M-step timed at full size."""
    from oracle import cpd_oracle as orc
    src, tgt = workload(points)
    s2 = float(orc.sigma2_init_exact(src, tgt))
    cols = min(cols, points)
    best = None
    for _ in range(repeats):
        t0 = time.perf_counter()
        ts = orc.apply_rigid(src, np.identity(3), np.zeros(3))
        es = orc.expectation_step(ts, tgt[:cols], s2, 0.0, n_global=points)
        t_e = time.perf_counter() - t0
        best = t_e if best is None else min(best, t_e)
    rng = np.random.default_rng(0)
    fake = orc.Estep(np.ones(points), rng.random(points) + 0.5, rng.random((points, 3)), float(points))
    fake = orc.Estep(fake.pt1, fake.p1, fake.px, float(fake.p1.sum()))
    t0 = time.perf_counter()
    orc.mstep_rigid(src, tgt, fake)
    t_m = time.perf_counter() - t0
    t_iter = best * (points / float(cols)) + t_m
    try:
        import threadpoolctl
        blas = max([p.get("num_threads", 1) for p in threadpoolctl.threadpool_info()] + [1])
    except Exception:
        blas = os.cpu_count()
    return {"value": 1.0 / t_iter, "unit": UNIT, "cores": int(blas), "host_cpus": os.cpu_count(), "kind": "port",
            "sample": "E-step on %d of %d target columns x all %d sources (%.2f s), extrapolated x%.0f; "
                      "+ full-size M-step (%.3f s); numpy/scipy port of probreg/cpd.py:71-88,160-192 "
                      "(~90%% single-threaded like the reference, BLAS threads = %d)"
                      % (cols, points, points, best, points / float(cols), t_m, blas),
            "sec_per_iter_extrapolated": t_iter, "ns_per_pair": best / (cols * float(points)) * 1e9}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cols = args.cpu_cols
    vals = []
    for _ in range(max(1, args.warmup > 0)):
        cpu_sample(args.points, min(cols, 200))
    last = None
    for _ in range(args.steps):
        last = cpu_sample(args.points, cols)
        vals.append(last["sec_per_iter_extrapolated"])
    t = float(np.mean(vals))
    last["value"] = 1.0 / t
    out = {"impl": "reference", "metric": METRIC, "value": 1.0 / t, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "strong",
           "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": {"workload": "rigid CPD, synthetic 3-D N=M=%d, sigma2 auto, w=0" % args.points,
                      "extrapolated_from_columns": cols},
           "cpu_baseline": last,
           "e2e": {"value": 1.0 / t, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0}
    emit(out)


# ---------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------
def run_ours(args):
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    import torch
    from probreg_b200 import _cabi, cpd, dist as pdist

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: probreg_b200 has no CPU path")
    torch.cuda.set_device(local_rank)
    comm = None
    if world > 1:
        import torch.distributed as tdist
        tdist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        comm = pdist.Communicator.from_torch(local_rank)

    def barrier():
        if world > 1:
            tdist.barrier()
        torch.cuda.synchronize()

    n = args.points
    src, tgt = workload(n)
    lo, hi = pdist.shard_bounds(n, rank, world)
    origin = tgt.mean(axis=0)

    # ---- device-resident arm -------------------------------------------------------------------
    h = _cabi.Handle(3, device=local_rank)
    if comm is not None:
        comm.attach(h)
    h.set_source(src)
    h.set_target(tgt[lo:hi], n_global=n, frame_origin=origin)
    s2 = h.sigma2_init()
    q0 = 1.0 + n * 3 * 0.5 * np.log(s2)

    def reset():
        h.set_state(_cabi.TF_RIGID, True, 0.0, np.identity(3), np.zeros(3), 1.0, s2, q0)

    reset()
    sampler = ClockSampler(world) if rank == 0 else None
    if sampler:
        sampler.start()
    for _ in range(max(args.warmup, 3)):
        h.em_step(read=False)
    h.sync()
    # keep the GPUs under the same load until the poller is up (its start-up must not fall into the timed
    # region); then rewind the EM state so that the timed iterations are the first ones of a registration
    t_up = time.perf_counter()
    ready = torch.zeros(1, device="cuda")
    while True:
        for _ in range(10):
            h.em_step(read=False)
        h.sync()
        ready[0] = 1.0 if (sampler is None or sampler.count() >= 2 * world or time.perf_counter() - t_up > 3.0) else 0.0
        if world > 1:
            tdist.broadcast(ready, src=0)
        if ready.item() > 0:
            break
    reset()
    for _ in range(3):
        h.em_step(read=False)
    h.sync()
    barrier()
    launches0 = h.launch_count()
    t_wall0 = time.perf_counter()
    for i in range(args.steps):
        h.flush_l2()
        h.event_record(2 * i)
        h.em_step(read=False)
        h.event_record(2 * i + 1)
    h.sync()
    barrier()
    t_wall1 = time.perf_counter()
    t_wall = t_wall1 - t_wall0
    launches = h.launch_count() - launches0
    clocks = sampler.stop(t_wall0, t_wall1) if sampler else None
    per_step = np.array([h.event_elapsed(2 * i, 2 * i + 1) for i in range(args.steps)])
    total_ms = float(per_step.sum())
    if world > 1:
        tt = torch.tensor([total_ms], dtype=torch.float64, device="cuda")
        tdist.all_reduce(tt, op=tdist.ReduceOp.MAX)
        total_ms = float(tt.item())
    ms_per_step = total_ms / args.steps
    value = 1e3 / ms_per_step
    final = h.em_step(read=True)            # the loop did real work: parameters moved towards the truth

    # ---- stage breakdown (profiling events, one sync per step; not part of `value`) -----------------
    h.set_profiling(True)
    stages = []
    for _ in range(5):
        h.flush_l2()
        h.em_step(read=False)
        stages.append(h.stage_times())
    h.set_profiling(False)
    st = np.median(np.array(stages), axis=0)
    names = ["pack", "pass1", "finalize1", "pass2", "finalize2", "moments_mstep"]
    t_estep_ms = float(st[1] + st[3])
    n_local = hi - lo

    # ---- end to end through the public API with host (pinned) buffers ---------------------------
    def pinned(a):
        t = torch.empty(a.shape, dtype=torch.float64, pin_memory=True)
        t.numpy()[...] = a
        return t.numpy()

    src_p, tgt_p = pinned(src), pinned(tgt)
    r = cpd.RigidCPD(src_p, device=local_rank, comm=comm)
    r.registration(tgt_p, maxiter=1, tol=-1.0)
    barrier()
    e2e_steps = max(3, min(args.steps, 10))
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        res = r.registration(tgt_p, maxiter=1, tol=-1.0)
    barrier()
    e2e_s = (time.perf_counter() - t0) / e2e_steps
    # amortised: one upload, K iterations, as a user's registration_cpd(maxiter=K) behaves
    t0 = time.perf_counter()
    r.registration(tgt_p, maxiter=args.steps, tol=-1.0)
    barrier()
    e2e_amort_s = (time.perf_counter() - t0) / args.steps
    if world > 1:
        tt = torch.tensor([e2e_s, e2e_amort_s], dtype=torch.float64, device="cuda")
        tdist.all_reduce(tt, op=tdist.ReduceOp.MAX)
        e2e_s, e2e_amort_s = float(tt[0].item()), float(tt[1].item())
    h2d = src.nbytes + tgt[lo:hi].nbytes
    d2h = 16 * 8

    if rank != 0:
        if world > 1:
            tdist.destroy_process_group()
        return

    # ---- rooflines -----------------------------------------------------------------------------
    probe = _cabi.microbench(local_rank)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    hbm_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback (B200_PROFILING.md)"
    flops = FLOP_PER_PAIR_ITER * float(n_local) * float(n)
    ach_tf = flops / (t_estep_ms * 1e-3) / 1e12
    nominal_tf = 2.0 * 128 * probe["sm_count"] * (clocks.get("sm_max_mhz") or 1965.0) * 1e6 / 1e12
    bytes_iter = hbm_bytes_per_iter(n_local, n)
    traffic, traffic_src = None, None
    try:   # dram__bytes_read.sum + dram__bytes_write.sum per launch of the two E-step kernels, from the committed ncu capture
        tj = json.load(open(os.path.join(ROOT, "profiles", "r1_traffic.json")))
        if world == 1 and n == 100000:
            traffic = sum(k["dram_read_bytes"] + k["dram_write_bytes"] for k in tj["kernels"].values())
            traffic_src = tj["source"]
    except Exception:
        pass
    roofline = {
        "bound": "fp32",
        "kernel": "pass1_kernel + pass2_kernel (fused E-step, never materialises P)",
        "achieved": ach_tf, "peak": probe["ffma_tflops"], "unit": "TFLOP/s", "frac": ach_tf / probe["ffma_tflops"],
        "peak_source": "FFMA issue-rate probe run in this process (cpd_microbench); nominal 2*128*SMs*clk = %.1f" % nominal_tf,
        "flop_per_pair": FLOP_PER_PAIR_ITER, "pairs_per_launch": float(n_local) * float(n),
        "instruction_ceiling_frac": ach_tf / (probe["ffma_tflops"] * (29.0 / (19.0 * 2.0))),
        "mufu_ex2_gops_probe": probe["mufu_ex2_gops"],
        "traffic": traffic,
        "traffic_source": traffic_src,
        "hbm": {"bound": "hbm", "achieved": bytes_iter / (t_estep_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                "frac": bytes_iter / (t_estep_ms * 1e-3) / 1e9 / hbm_peak, "peak_source": hbm_src,
                "algorithmic_bytes": bytes_iter,
                "materialised_equiv_gbs": 8.0 * n_local * n / (t_estep_ms * 1e-3) / 1e9,
                "materialised_equiv_frac": 8.0 * n_local * n / (t_estep_ms * 1e-3) / 1e9 / hbm_peak},
    }
    cpu = cpu_sample(n, args.cpu_cols) if world == 1 and not args.no_cpu else None
    extras = run_extras() if world == 1 and not args.no_extras and n == 100000 else None
    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32", "dtype_note": "pair arithmetic f32 (packed f32x2); every sum beyond 64 terms, the moments and the M-step f64",
        "data": "synthetic",
        "config": {"workload": "rigid CPD, synthetic 3-D N=M=%d, sigma2 auto, w=0, update_scale" % n,
                   "parallelism": "target-sharded x%d, sources replicated, one 32-double all-reduce per iteration (%s)"
                                  % (world, "fused into the M-step kernel over NVLink peer memory" if (comm is not None and comm.use_p2p)
                                     else ("ncclAllReduce" if world > 1 else "none needed")),
                   "l2": "flushed (256 MiB memset) between timed iterations, outside the event pairs",
                   "timing": "CUDA event pair per iteration on the library stream, summed, max over ranks"},
        "wall_ms_per_step_incl_flush": t_wall * 1e3 / args.steps,
        "e2e": {"value": 1.0 / e2e_s, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                "what": "RigidCPD.registration(target, maxiter=1) per step: H2D both clouds (pinned), sigma2 init, "
                        "1 EM iteration, D2H MstepResult",
                "amortised_value": 1.0 / e2e_amort_s,
                "amortised_what": "registration(maxiter=%d): one upload, per-iteration D2H of the MstepResult" % args.steps},
        "gpu_launches": int(launches),
        "stage_ms": dict(zip(names, [float(x) for x in st])),
        "roofline": roofline,
        "cpu_baseline": cpu,
        "clocks": clocks,
        "probe": probe,
        "result_check": {"sigma2_after_run": final[3], "scale": final[2]},
        "extras": extras,
    }
    emit(out)
    if world > 1:
        tdist.destroy_process_group()


def run_extras():
    """Side measurements (BASELINE configs 3 and 5, first-hardware-run probes of the newest paths) in a SUBPROCESS, after the bench
    line's own numbers are final: whatever happens there -- an exception, a CUDA fault, a time-out -- costs the bench line
    nothing but this key.  tools/bench_extras.py prints one JSON object."""
    import subprocess

    try:
        r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "bench_extras.py")], capture_output=True, text=True, timeout=420,
                           cwd=ROOT)
        lines = [l for l in r.stdout.splitlines() if l.strip().startswith("{")]
        if r.returncode != 0 or not lines:
            return {"error": "exit %d: %s" % (r.returncode, (r.stderr or r.stdout)[-400:])}
        return json.loads(lines[-1])
    except Exception as e:                      # noqa: BLE001
        return {"error": "%s: %s" % (type(e).__name__, str(e)[:300])}


def emit(obj):
    """The ONE JSON line goes to the real stdout; everything else this process (or NCCL, which
    prints its version banner to fd 1) writes during the run has been diverted to stderr."""
    os.write(_REAL_STDOUT, (json.dumps(obj) + "\n").encode())


_REAL_STDOUT = os.dup(1)
os.dup2(2, 1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--points", type=int, default=100000)
    ap.add_argument("--cpu-cols", type=int, default=2000, help="target columns in the CPU sample")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-extras", action="store_true", help="skip the side measurements of tools/bench_extras.py (key \"extras\")")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
