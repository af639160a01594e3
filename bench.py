#!/usr/bin/env python
"""bench.py -- stereo frames/s (match + triangulate + track) on N B200s vs the host CPU.

Metric (BASELINE.json): stereo frames/sec for the hot path match -> DLT triangulate ->
reproject -> LSAP -> rows -> frame-to-frame tracking assignment (+ chain pass).

Workload at N=1: BASELINE.json configs[1] -- 1,000 synthetic frames x 8 colours x 32 rods
(P = 8,000 matching problems, 7,992 tracking problems per step).  N>1: every rank runs that
workload on its own contiguous frame range (weak scaling); per step the ranks exchange the
1-frame tracking halo, all-gather the chain-pass boundary maps and gather rows + track ids to
rank 0 over NCCL.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Prints ONE JSON line on rank 0 (see the contract in the task description):
  value    whole-job frames/s, inputs resident in HBM, CUDA-event timed, max over ranks
  e2e      same metric through the host-buffer C-ABI entry point (ptk_reconstruct_host) with
           pinned host inputs/outputs, H2D + D2H inside every call; two caller threads in flight
           (the call is re-entrant; the reference's own callers use one thread per colour), the
           single synchronous caller's figure beside it
  roofline dominant kernel (K1 k_cost): algorithmic FP64 flop (as built; the SURVEY model of the
           reference's algorithm beside it) / event-timed duration vs the FP64 FMA peak measured
           in this run (MEASURED_PEAKS.json has no FP64 figure), plus its algorithmic HBM GB/s vs
           MEASURED_PEAKS.json hbm_gbs
  cpu_baseline  the oracle port (cv2 + scipy, what the reference calls) on one host core on a
           bounded sample of the same workload
`--impl reference` times the reference's CPU path (oracle port: cv2.triangulatePoints +
scipy.linear_sum_assignment per (frame, colour), then the tracking assignment) on all host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

FRAMES, COLORS, RODS = 1000, 8, 32
WORKLOAD = "match_csv_complex+tracking_global_assignment: 1000 frames x 8 colours x 32 rods (BASELINE configs[1])"
METRIC = "stereo frames/sec (match+triangulate+track)"
# Algorithmic FP64 work of K1 per rod pair (4 end-point combinations), FMA = 2 flop:
#  * as built (canonical rigs: camera-2 records + operators on the camera-1 point, two squarings of a 4x4
#    matrix and three products from the start vector, homogeneous reprojection): 141.3 DFMA + 61.2 DMUL + 18.1 DADD executed per
#    combination (ncu counters, profiles/r02r_k_cost_ncu.txt) = 362.0 flop -> 1 448 per pair (round 1:
#    2 400, start of round 2: 2 244 -- the kernel got faster mostly by doing less, which a flop-per-second
#    fraction only half shows; roofline.fp64_pipe_busy is the pipe's own utilisation).  This is what
#    roofline.achieved / frac are computed from: the kernel's own work against the FP64 FMA peak.
#  * SURVEY.md 8d / App. H model of the REFERENCE's algorithm (OpenCV's 4x4 Jacobi SVD per point):
#    245 + 80*R + 66*S flop per combination with R = 20.5 rotations, S = 4.9 sweeps -> 2 210 per
#    combination, 8 840 per pair.  Reported beside it as roofline.reference_algorithm (a ratio above
#    1 there means: faster than the reference's algorithm could run at FP64 peak).
FLOP_PER_PAIR = 1448.0
FP64_PIPE_INSTR_PER_COMBINATION = 224.6  # 141.3 DFMA + 61.2 DMUL + 18.1 DADD + 4 DSETP (same capture)
FLOP_PER_PAIR_REFERENCE_ALGORITHM = 8840.0
N_BATCHES = 4  # distinct synthetic input batches the timed steps rotate over


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--frames", type=int, default=FRAMES, help=argparse.SUPPRESS)
    ap.add_argument("--rods", type=int, default=RODS, help=argparse.SUPPRESS)
    ap.add_argument("--colors", type=int, default=COLORS, help=argparse.SUPPRESS)
    ap.add_argument("--no-cpu-baseline", action="store_true", help=argparse.SUPPRESS)
    ap.add_argument("--config4", action="store_true", help=argparse.SUPPRESS)  # extra pass: BASELINE configs[3] (default at 8 GPUs)
    ap.add_argument("--no-sustained", action="store_true", help=argparse.SUPPRESS)
    return ap.parse_args()


# ------------------------------------------------------------------------------------------
# CPU side (oracle port) -- the only place bench.py executes oracle/
# ------------------------------------------------------------------------------------------
def _cpu_worker(args):
    """Reference CPU path on a frame range: per colour, per frame match (cv2 + scipy), then the
    tracking assignment per colour.  Returns frames processed."""
    import cv2

    cv2.setNumThreads(1)
    from scipy.spatial.transform import Rotation

    from oracle import port
    from particletracking_b200 import synthetic as syn

    cam1, cam2, seed = args
    F, C, N = cam1.shape[:3]
    cal, trf = syn.synthetic_calibration(), syn.synthetic_transformation()
    P1, P2, r1, r2, t1, t2 = syn.projection_matrices(cal)
    rot = Rotation.from_matrix(trf["rotation"])
    for c in range(C):
        ends = np.zeros((F, N, 2, 3))
        for f in range(F):
            r = port.match_arrays(cam1[f, c], cam2[f, c], cal, P1, P2, rot, trf["translation"], r1, r2, t1, t2)
            ends[f] = r["rows"][:, :6].reshape(N, 2, 3)
        if F > 1:
            port.tracking_arrays(ends[:, :, 0], ends[:, :, 1])
    return F


def cpu_port_fps(n_frames: int, rods: int, colors: int, cores: int, pool=None):
    """frames/s of the oracle port on `cores` processes, `n_frames` frames of the workload."""
    from particletracking_b200 import synthetic as syn

    data = syn.make_stereo(n_frames, colors, rods, seed=99)
    if cores <= 1:
        t0 = time.perf_counter()
        _cpu_worker((data.cam1, data.cam2, 0))
        return n_frames / (time.perf_counter() - t0)
    per = [np.array_split(np.arange(n_frames), cores)[i] for i in range(cores)]
    jobs = []
    for idx in per:
        if len(idx) == 0:
            continue
        lo, hi = idx[0], min(idx[-1] + 2, n_frames)  # 1-frame overlap so every pair is tracked once
        jobs.append((data.cam1[lo:hi], data.cam2[lo:hi], 0))
    t0 = time.perf_counter()
    pool.map(_cpu_worker, jobs)
    return n_frames / (time.perf_counter() - t0)


def run_reference(args):
    """`--impl reference`: the reference's CPU path on all host cores (rank 0 only)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp

    cores = len(os.sched_getaffinity(0))
    frames_per_step = max(cores, 8) * 2
    ctx = mp.get_context("fork")
    with ctx.Pool(cores) as pool:
        for _ in range(args.warmup):
            cpu_port_fps(frames_per_step, args.rods, args.colors, cores, pool)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            cpu_port_fps(frames_per_step, args.rods, args.colors, cores, pool)
        dt = time.perf_counter() - t0
    fps = frames_per_step * args.steps / dt
    sample = f"{frames_per_step} frames x {args.colors} colours x {args.rods} rods per step, frame ranges over {cores} processes"
    line = {
        "impl": "reference", "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample": sample},
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock and throttle reasons DURING the timed region.  NVML is read in-process every 10 ms
    (microseconds per call); the `nvidia-smi -lms` child used before held the driver's lock for
    milliseconds per poll, which stalled rank 0's CUDA calls and made short multi-GPU steps
    host-bound in some runs (5.8 vs 1.7 ms per step at 2 GPUs).  nvidia-smi remains the fallback."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")
    NVML_REASONS = ((0x8, "hw_slowdown"), (0x40, "hw_thermal_slowdown"), (0x20, "sw_thermal_slowdown"), (0x4, "sw_power_cap"))

    def __init__(self, gpu_index: int):
        self.idx = gpu_index
        self.lines = []    # nvidia-smi fallback: (t, csv line)
        self.samples = []  # NVML: (t, sm_mhz, reasons bitmask)
        self.proc = None
        self.nvml = None
        self.max_mhz = None
        self._stop = threading.Event()

    def start(self):
        try:
            import pynvml

            pynvml.nvmlInit()
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(self.idx)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            self.nvml = pynvml
            threading.Thread(target=self._poll, daemon=True).start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50", "-i", str(self.idx)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _poll(self):
        n = self.nvml
        reasons = getattr(n, "nvmlDeviceGetCurrentClocksEventReasons", None) or n.nvmlDeviceGetCurrentClocksThrottleReasons
        while not self._stop.is_set():
            try:
                self.samples.append((time.perf_counter(), float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)),
                                     int(reasons(self.handle))))
            except Exception:
                pass
            time.sleep(0.01)

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append((time.perf_counter(), ln.strip()))

    def stop(self, t0=None, t1=None):
        if self.nvml is not None:
            time.sleep(0.03)
            self._stop.set()
            sm, reasons = [], set()
            for ts, mhz, bits in self.samples:
                if t0 is not None and not (t0 - 0.02 <= ts <= t1 + 0.03):
                    continue
                sm.append(mhz)
                for bit, name in self.NVML_REASONS:
                    if bits & bit:
                        reasons.add(name)
            return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(reasons),
                    "samples": len(sm), "source": "nvml"}
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        for ts, ln in self.lines:
            if t0 is not None and not (t0 - 0.05 <= ts <= t1 + 0.15):
                continue
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx = float(f[2])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm), "source": "nvidia-smi"}


# ------------------------------------------------------------------------------------------
# GPU side
# ------------------------------------------------------------------------------------------
def run_b200(args):
    import torch
    import torch.distributed as dist

    from particletracking_b200 import distributed as pdist
    from particletracking_b200 import engine, rig as rigmod, synthetic as syn

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a GPU (there is no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # NCCL_DEBUG is left exactly as the environment sets it (the driver counts ranks from its INFO lines)
        dist.init_process_group("nccl", device_id=dev)
    F, C, N = args.frames, args.colors, args.rods
    K, W = args.steps, max(args.warmup, 3)

    # ---- synthetic inputs: N_BATCHES distinct experiments per rank, resident in HBM.  N > 1: every rank but
    # the last is also handed the first frame of the next rank's range (the halo frame it re-computes
    # instead of receiving it): F + 1 frames of detections.
    rig = None
    batches, host_batches = [], []
    halo = 1 if (world > 1 and rank < world - 1) else 0
    for b in range(N_BATCHES):
        data = syn.make_stereo(F, C, N, seed=1000 * (rank + 1) + b)
        if rig is None:
            rig = rigmod.rig_from_dicts(data.calibration, data.transformation)
        arrs = list(data.problems())
        if halo:
            nxt = syn.make_stereo(F, C, N, seed=1000 * (rank + 2) + b).problems()
            arrs = [np.concatenate([a, x[:C]], axis=0) for a, x in zip(arrs, nxt)]
        batches.append(tuple(torch.from_numpy(np.ascontiguousarray(x)).to(dev) for x in arrs))
        if b < 2:
            host_batches.append(tuple(torch.from_numpy(np.ascontiguousarray(x)).pin_memory()
                                      for x in (data.cam1, data.cam2, data.n1, data.n2)))
    sharded = pdist.ShardedReconstructor(rig, F, C, N, dev) if world > 1 else None
    pipe_dev = engine.DevicePipeline(rig, F, C, N, dev, track=True) if world == 1 else None

    def step(i):
        c1, c2, n1, n2 = batches[i % N_BATCHES]
        if sharded is not None:
            return sharded.step(c1, c2, n1, n2)
        return pipe_dev.run(c1, c2, n1, n2)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    fp64_peak = engine.measure_fp64_peak(20000)  # also warms the clocks
    for i in range(W):
        out = step(i)
    # Settling.  The timed region enqueues K steps without synchronising, so several steps' worth of
    # output tensors and asynchronous result gathers are in flight at once; until torch's caching
    # allocator has grown its pool to that depth, a step now and then pays for a cudaMalloc (and on a
    # fresh box NCCL keeps paying host-side set-up costs for dozens of steps).  Measured at 2 and 4
    # GPUs: the same build gave 1.74 or 3.05-5.77 ms per step depending on whether that happened
    # inside the timed region.  So the warm-up runs in the timed region's own pattern -- untimed
    # bursts of K unsynchronised steps -- until two consecutive bursts agree within 10 % and the last
    # one is within 10 % of the best seen; at most 6 s.  The exit decision is collective (a rank-local
    # clock would leave the other ranks inside a collective: one 2-GPU run hung that way).  The timed
    # region below is still exactly K steps.
    # Python's cyclic collector stays out of the measured regions: a generation-2 pass over this
    # process's heap (torch + numpy + pandas) is tens of milliseconds, the timed region of 20 steps
    # is 35 ms, and with N ranks coupled by a halo exchange per step one rank's pause is everyone's.
    import gc

    gc.collect()
    gc.freeze()
    gc.disable()
    extra, hist, t_settle0 = 0, [], time.perf_counter()
    burst = max(K, 10)
    while True:
        t_s = time.perf_counter()
        for j in range(burst):
            out = step(W + extra + j)
        if sharded is not None:
            sharded.finish()
        torch.cuda.synchronize()
        hist.append(time.perf_counter() - t_s)
        extra += burst
        stable = len(hist) >= 2 and max(hist[-2:]) <= 1.10 * min(hist[-2:]) and hist[-1] <= 1.10 * min(hist)
        timed_out = time.perf_counter() - t_settle0 >= 6.0
        if world > 1:
            flag = torch.tensor([1.0 if stable else 0.0, 0.0 if timed_out else 1.0], device=dev)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            flag = flag.tolist()
            stable, timed_out = flag[0] == 1.0, flag[1] == 0.0
        if stable or timed_out:
            break
    W_total = W + extra
    if sharded is not None:
        sharded.finish()
    barrier()
    if sharded is None:
        assert int((out.status != 0).sum().item()) == 0, "matching reported invalid problems"
        assert int((out.track_status != 0).sum().item()) == 0, "tracking reported invalid problems"
    else:
        assert int((out["status"] != 0).sum().item()) == 0, "matching reported invalid problems"
        if rank == 0:  # the gathered blocks of every rank: all problems solved, ids are permutations in every frame
            assert all(int((x != 0).sum().item()) == 0 for x in out["status_all"] + out["track_status_all"])
            tid_all = torch.cat(list(out["track_id_all"]), dim=1)
            assert tuple(tid_all.shape) == (C, world * F, N)
            assert bool((tid_all.sort(dim=-1).values == torch.arange(N, device=dev, dtype=tid_all.dtype)).all())

    clocks = ClockSampler(local)
    dbg = os.environ.get("PTK_BENCH_DEBUG", "")  # experiment switches: "nosampler", "notiming", "noe2e"
    if rank == 0 and "nosampler" not in dbg:  # one sampler per job: nvidia-smi polling from every rank perturbs the driver
        clocks.start()
    time.sleep(0.3)
    if sharded is not None and getattr(sharded, "_prof", None) is not None:
        if rank == 0:
            print("step() host profile, warm-up + settle, ms per step: " +
                  json.dumps({k: round(v / max(sharded._step_idx, 1) * 1e3, 3) for k, v in sharded._prof.items()}), file=sys.stderr)
        sharded._prof.clear()
        prof_step0 = sharded._step_idx
    # CUDA events inside the timed region only at N = 1 (one library call per step) and only around the
    # dominant kernel (K1, the roofline's duration): event pairs around all ten launches of a step cost the
    # region 4 % (1.083 vs 1.035 ms per step).  The other kernels' breakdown comes from a second, untimed
    # pass of K steps right after the timed region -- at N > 1 the whole breakdown does (graph replays are
    # not event-timed, and event pairs around every launch together with the NVML sampler made rank 0's
    # host 5x slower in half of the 4-GPU runs of round 1).
    timing_in_region = world == 1 and "notiming" not in dbg
    engine.kernel_timing(2 if timing_in_region else False)
    engine.kernel_timing_read(reset=True)
    engine.launch_count(reset=True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    seg0 = torch.cuda.memory_stats(dev).get("segment.all.allocated", 0)
    t_wall0 = time.perf_counter()
    e0.record()
    for i in range(K):
        step(W_total + i)
    if sharded is not None:
        sharded.finish()  # the last steps' result gathers are part of the timed region
    e1.record()
    t_enq = time.perf_counter() - t_wall0  # host time to enqueue K steps (no synchronisation inside)
    barrier()
    t_wall1 = time.perf_counter()
    ms = e0.elapsed_time(e1)
    seg_added = torch.cuda.memory_stats(dev).get("segment.all.allocated", 0) - seg0  # cudaMallocs inside the timed region
    launches = engine.launch_count()
    clk = clocks.stop(t_wall0, t_wall1)
    timing_region = engine.kernel_timing_read(reset=True) if timing_in_region else None
    timing = timing_region
    if "notiming" not in dbg:
        engine.kernel_timing(True)  # graph replays are not event-timed: this pass launches eagerly
        engine.kernel_timing_read(reset=True)
        for i in range(K):
            step(W_total + K + i)
        if sharded is not None:
            sharded.finish()
        barrier()
        timing = engine.kernel_timing_read(reset=True)
    if timing_region is not None:
        timing = dict(timing, cost=timing_region["cost"])  # K1: the figure measured inside the timed region
    k1_ms, k1_n = timing["cost"]
    engine.kernel_timing(False)
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    value = world * F * K / (ms * 1e-3)

    # ---- sustained figure: the same step back to back for >= 2 s (the K-step region above is a
    # ~30 ms burst), with its own clock record; not the headline, reported beside it
    sustained = None
    if not args.no_sustained and "nosustained" not in dbg:
        n_sus = int(min(max(2000.0 / (ms / K) * 1.05, 4 * K), 20000))
        clocks_s = ClockSampler(local)
        if rank == 0 and "nosampler" not in dbg:
            clocks_s.start()
        barrier()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        tw0 = time.perf_counter()
        s0.record()
        for i in range(n_sus):
            step(W_total + 2 * K + i)
            if i % 64 == 63:  # keep the host at most 64 steps ahead of the GPU
                torch.cuda.current_stream().synchronize()
        if sharded is not None:
            sharded.finish()
        s1.record()
        barrier()
        tw1 = time.perf_counter()
        ms_s = s0.elapsed_time(s1)
        if world > 1:
            t = torch.tensor([ms_s], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms_s = float(t.item())
        sustained = {"seconds": ms_s * 1e-3, "steps": n_sus, "ms_per_step": ms_s / n_sus,
                     "value": world * F * n_sus / (ms_s * 1e-3), "unit": "frames/s", "clocks": clocks_s.stop(tw0, tw1)}

    # ---- BASELINE configs[3] (50 000 frames x 8 colours x 64 rods over the GPUs of the box): one pass
    config4 = None
    if args.config4 or world == 8:
        config4 = run_config4(world, rank, dev, rig, barrier)

    # ---- e2e: host buffers through the host entry points (H2D + D2H inside every call), every rank its own shard.
    # Two result forms: the compact one (ptk_reconstruct_host_compact: what the device computed -- 3D end points,
    # pairing verdict, costs, assignments, tracks; the 18 reference columns follow from it and the caller's own
    # inputs, engine.expand_rows) and the full 18-column rows (ptk_reconstruct_host).  Each with one synchronous
    # caller and with two callers in flight (the reference's GUI calls the path from one thread per colour,
    # SURVEY 8b): two host pipelines, two threads -- the head of one call hides under the tail of the other.
    Ke = 1 if "noe2e" in dbg else max(3, min(K, 10))
    P = F * C
    h2d = P * N * 4 * 8 * 2 + P * 4 * 2
    small = P * N * 8 + 2 * P * N * 4 + 2 * P * 4 + C * (F - 1) * (N * 4 + 8 + 4) + C * F * N * 4
    d2h_full = P * N * 18 * 8 + small
    d2h_compact = P * N * 6 * 8 + P * N + small

    def e2e_pair(compact):
        pa, pb = engine.HostPipeline(local), engine.HostPipeline(local)
        for i in range(2):
            pa.reconstruct(rig, *host_batches[i % 2], track=True, compact=compact)
        pb.reconstruct(rig, *host_batches[1], track=True, compact=compact)
        barrier()
        t0 = time.perf_counter()
        for i in range(Ke):
            pa.reconstruct(rig, *host_batches[i % 2], track=True, compact=compact)
        torch.cuda.synchronize()
        one_s = time.perf_counter() - t0

        def _caller(pp, idx):
            for _ in range(Ke):
                pp.reconstruct(rig, *host_batches[idx], track=True, compact=compact)

        barrier()
        t0 = time.perf_counter()
        th = [threading.Thread(target=_caller, args=(pp, i)) for i, pp in enumerate((pa, pb))]
        for t_ in th:
            t_.start()
        for t_ in th:
            t_.join()
        torch.cuda.synchronize()
        two_s = time.perf_counter() - t0
        last = pa.reconstruct(rig, *host_batches[0], track=True, compact=compact)
        assert int((last["status"] != 0).sum()) == 0
        expand_ms = None
        if compact and rank == 0:  # what rebuilding the 18 columns costs on this host (outside every timed region)
            engine.expand_rows(last, host_batches[0][0], host_batches[0][1])
            t0 = time.perf_counter()
            engine.expand_rows(last, host_batches[0][0], host_batches[0][1])
            expand_ms = (time.perf_counter() - t0) * 1e3
        pa.close()
        pb.close()
        if world > 1:
            t = torch.tensor([one_s, two_s], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            one_s, two_s = (float(x) for x in t.tolist())
        return world * F * Ke / one_s, world * F * 2 * Ke / two_s, expand_ms

    c_one, c_two, expand_ms = e2e_pair(True)
    f_one, f_two, _ = e2e_pair(False)

    # ---- roofline of the dominant kernel (K1)
    pairs_per_launch = P * N * N / max(k1_n / K, 1) if k1_n else P * N * N
    k1_avg_ms = k1_ms / k1_n if k1_n else float("nan")
    flops_per_launch = FLOP_PER_PAIR * pairs_per_launch
    achieved_tf = flops_per_launch / (k1_avg_ms * 1e-3) / 1e12
    sm_clock_hz = (clk.get("sm_mhz") or 0.0) * 1e6  # median SM clock sampled during the timed region
    n_schedulers = 4 * torch.cuda.get_device_properties(local).multi_processor_count
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    bytes_per_launch = pairs_per_launch * (8 + 1) + (pairs_per_launch / N) * 2 * 4 * 8 * 2  # cost+choice out, und+orig rods in
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "k_cost_traffic.json"))).get("dram_bytes_per_launch")
    except Exception:
        pass
    ref_tf = FLOP_PER_PAIR_REFERENCE_ALGORITHM * pairs_per_launch / (k1_avg_ms * 1e-3) / 1e12
    roofline = {
        "kernel": "k_cost_rows<simple> (K1, canonical rig: DLT null vector by inverse power iteration on per-end-point "
                  "records, reprojection; 1 thread per end-point combination, row-block CTAs)",
        "bound": "fp64", "achieved": achieved_tf, "peak": fp64_peak / 1e12, "unit": "TFLOP/s",
        "frac": achieved_tf / (fp64_peak / 1e12),
        "peak_source": "measured in this run: dependency-free DFMA loop on all SMs (ptk_measure_fp64_peak); "
                       "MEASURED_PEAKS.json has no FP64 figure",
        "algorithmic_flop_per_pair": FLOP_PER_PAIR, "pairs_per_launch": pairs_per_launch,
        "flop_model": "as built: 141.3 DFMA + 61.2 DMUL + 18.1 DADD per combination (ncu, profiles/r02r_k_cost_ncu.txt), FMA = 2",
        # utilisation of the FP64 pipe itself: FP64-pipe warp instructions x 2 issue cycles (16 lanes per scheduler) over
        # the kernel's scheduler-cycles; ncu's sm__pipe_fp64_cycles_active of the same build: 78.2 %
        "fp64_pipe_busy": FP64_PIPE_INSTR_PER_COMBINATION * 2.0 * (pairs_per_launch * 4 / 32) /
                          (k1_avg_ms * 1e-3 * sm_clock_hz * n_schedulers) if sm_clock_hz else None,
        "reference_algorithm": {"flop_per_pair": FLOP_PER_PAIR_REFERENCE_ALGORITHM, "achieved": ref_tf,
                                "frac": ref_tf / (fp64_peak / 1e12),
                                "note": "SURVEY 8d model of OpenCV's Jacobi SVD; > 1 = algorithmic saving"},
        "kernel_ms": k1_avg_ms, "kernel_share_of_step": (k1_ms / K) / (ms / K),
        # whole step against the same peak: FP64 flop of K1 + K3 (the re-evaluated matched pairs) per step / step time
        "step_frac": FLOP_PER_PAIR * (P * N * N + P * N) / (ms / K * 1e-3) / fp64_peak,
        "hbm": {"achieved": bytes_per_launch / (k1_avg_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                "frac": bytes_per_launch / (k1_avg_ms * 1e-3) / 1e9 / hbm_peak,
                "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback"},
        "traffic": traffic,
    }

    if sharded is not None and getattr(sharded, "_prof", None) and rank == 0:
        print("step() host profile, timed region (+ e2e steps), ms per step: " +
              json.dumps({k: round(v / max(sharded._step_idx - prof_step0, 1) * 1e3, 3) for k, v in sharded._prof.items()}), file=sys.stderr)
    gc.enable()
    line = {
        "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": K, "warmup": W, "warmup_settle_steps": extra,
        "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "frames_per_gpu": F, "colours": C, "rods": N, "problems_per_step": P,
                   "l2": f"inputs rotate over {N_BATCHES} distinct batches; per-step working set "
                         f"{(P * N * N * 8 * 2 + P * N * 18 * 8) / 1e6:.0f} MB > 126 MB L2",
                   "parallelism": f"frames sharded over {world} GPU(s)" + ("; halo frame re-computed locally, one library call (CUDA graph) + one "
                                                                               "NCCL gather of the result block per step, root-side re-base of track ids" if world > 1 else "")},
        "e2e": {"value": c_two, "unit": "frames/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h_compact,
                "steps": 2 * Ke,
                "api": "ptk_reconstruct_host_compact (HostPipeline.reconstruct(compact=True)) from pinned host buffers, H2D + D2H inside "
                       "every call; two caller threads in flight, one HostPipeline each (the reference's callers run one thread per colour)",
                "result": "compact: 3D end points [F,C,N,6], straight/crossed verdict, match costs, row/col assignments, track "
                          "assignments, totals, ids, status words; the reference's 18 columns are rebuilt bit for bit from these and the "
                          "caller's own inputs by ptk_expand_rows_host (not in the timed region; its cost on this host: expand_rows_ms)",
                "one_synchronous_caller": {"value": c_one, "unit": "frames/s", "steps": Ke},
                "expand_rows_ms": expand_ms,
                "full_rows": {"value": f_two, "unit": "frames/s", "d2h_bytes_per_step": d2h_full,
                              "api": "ptk_reconstruct_host: the 18-column rows travel (the round-1 e2e figure); device-to-host writes "
                                     "bound it on a multi-GPU host (profiles/r02g_copy_probe_8gpu.json)",
                              "one_synchronous_caller": {"value": f_one, "unit": "frames/s", "steps": Ke}}},
        "gpu_launches": launches,
        "roofline": roofline,
        "clocks": clk,
        "kernel_ms_per_step": {k: v[0] / K for k, v in timing.items() if v[1]},
        "wall_s": t_wall1 - t_wall0,
        "host_enqueue_ms_per_step": t_enq / K * 1e3,
        "cuda_mallocs_in_timed_region": int(seg_added),
        "kernel_timing": "K1: CUDA events around its launches inside the timed region; the other kernels: a second pass of K steps right after it" if timing_in_region
                         else "separate untimed pass of K steps after the timed region (eager launches; the timed region replays CUDA graphs)",
        "sustained": sustained,
        "tracking_parity": "col_ind / track ids are bit-exact given identical 3D input (the reference's tracking cost is separable up "
                           "to rounding, so SciPy's choice hangs on the last bits of the rows; DESIGN.md 4)",
        "reference_arm": "--impl reference runs the oracle port (oracle/port.py: the reference's Python restated around the same "
                         "cv2 / scipy calls), about 2.3x faster per core than the stock reference (BASELINE.md 2): ratios against it "
                         "understate the speed-up over the real reference by that factor",
    }
    if config4 is not None:
        line["config4"] = config4

    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        sample_frames = 24
        fps1 = cpu_port_fps(sample_frames, N, C, 1)
        line["cpu_baseline"] = {"value": fps1, "unit": "frames/s", "cores": 1, "kind": "port",
                                "sample": f"{sample_frames} frames x {C} colours x {N} rods (oracle/port.py: cv2.triangulatePoints "
                                          f"+ projectPoints + scipy LSAP per (frame, colour), then tracking), one process, cv2 threads = 1"}
        try:
            from oracle import c_oracle as co

            data = syn.make_stereo(200, C, N, seed=7)
            c1, c2, n1, n2 = data.problems()
            nt = co.num_threads()
            t0 = time.perf_counter()
            o = co.match_batch(rig, c1, c2, n1, n2, nthreads=nt)
            ends = o["rows"][..., :6].reshape(200, C, N, 2, 3).transpose(1, 0, 2, 3, 4).copy()
            co.track_batch(ends, nthreads=nt)
            line["cpu_c_oracle"] = {"value": 200 / (time.perf_counter() - t0), "unit": "frames/s", "cores": nt,
                                    "kind": "port", "sample": "200 frames, oracle/ptk_oracle.c (plain C, OpenMP)"}
        except Exception as e:  # the C oracle is optional for the bench
            line["cpu_c_oracle"] = {"error": str(e)}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def run_config4(world, rank, dev, rig, barrier):
    """BASELINE configs[3]: matching + frame-to-frame tracking over 50 000 frames x 8 colours x 64 rods,
    sharded over the ranks (50 000 / world frames each), one pass after one warm-up pass; max over ranks."""
    import torch
    import torch.distributed as dist

    from particletracking_b200 import distributed as pdist
    from particletracking_b200 import engine, synthetic as syn

    Ftot, C, N = 50000, 8, 64
    f0, f1 = pdist.shard_frames(Ftot, world, rank)
    i0, i1 = pdist.shard_input_range(Ftot, world, rank)
    # synthetic detections for this rank's range (+ the halo frame it re-computes)
    data = syn.make_stereo(i1 - i0, C, N, seed=4000 + rank, first_frame=i0)
    arrs = tuple(torch.from_numpy(np.ascontiguousarray(x)).to(dev) for x in data.problems())
    del data
    F = f1 - f0
    if world > 1:
        rec = pdist.ShardedReconstructor(rig, F, C, N, dev, n_frames_total=Ftot, use_graph=False)
        run = lambda: rec.step(*arrs)
        fin = rec.finish
    else:
        pipe = engine.DevicePipeline(rig, F, C, N, dev, track=True)
        run = lambda: pipe.run(*arrs)
        fin = lambda: None
    out = run()
    fin()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    out = run()
    fin()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    st = out["status"] if world > 1 else out.status
    tst = out["track_status"] if world > 1 else out.track_status
    ok = int((st != 0).sum().item()) == 0 and int((tst != 0).sum().item()) == 0
    if world > 1:
        t = torch.tensor([ms, 0.0 if ok else 1.0], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, ok = float(t[0].item()), float(t[1].item()) == 0.0
        if rank == 0:
            ok = ok and all(int((x != 0).sum().item()) == 0 for x in out["status_all"] + out["track_status_all"])
    return {"workload": "BASELINE configs[3]: 50 000 frames x 8 colours x 64 rods, match + track + chain", "frames": Ftot,
            "frames_per_gpu": F, "ms": ms, "frames_per_s": Ftot / (ms * 1e-3), "status_ok": bool(ok),
            "note": "one pass, device-resident inputs, final gather to rank 0 included" if world > 1 else "one pass, device-resident inputs"}


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)
