"""bench.py -- candidate iLQR solves/s of the i2LQR control solve (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--batch B]

A "step" is one pass of the hot path over one batch of synthetic input: B independent candidate iLQR solves
per GPU (BASELINE config 5: 2^20 flat candidates, horizon 6, per-candidate obstacle near the path, seed
20260101 + rank).  Weak scaling: every rank owns its own B candidates, there is no data-path collective.

One JSON line on rank 0:
  value      whole-job solves/s with inputs resident in HBM (CUDA events on the launching stream, max over ranks)
  e2e        the same metric through the C-ABI host call (pinned host buffers, H2D + D2H inside the timed region)
  roofline   FP64 FMA-pipe: executed iLQR iterations (device counter) x (927 N + 177) algorithmic flops
              / kernel time, against the DFMA peak measured on this GPU in the same run
  cpu_baseline  the C oracle port (kind "port") on all host cores over a bounded sample of the same workload
  latency    p50 / p90 of one full controller step (BASELINE config 2 shape: K=64, L=8) through the C ABI

--impl reference times the CPU implementation alone (the reference is pure Python and cannot travel to the
GPU box; the C oracle restates it -- kind "port"), same metric / config, rank 0 only.
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

# dram bytes of the dominant launch (pass 0 over 2^20 candidates) from profiles/r02_solve_kernel_v8_lean_pass0_summary.txt
NCU_TRAFFIC_BYTES = 135843072 + 307914752
# what the kernel EXECUTES (it exploits the sparsity of A / B / Vxx, so fewer flops than the dense algorithmic count) and
# how busy the FP64 pipe is, from the committed ncu capture of the same launch -- a static annotation of the line, the live
# numbers are `achieved` / `frac`
NCU_EXECUTED = {
    "source": "profiles/r02_solve_kernel_v8_lean_pass0_summary.txt (ncu --set full, pass-0 launch over 2^20 candidates, 2.32 ms)",
    "pipe_fp64_busy": 0.632, "executed_tflops": 16.6, "executed_frac_of_measured_peak": 0.49,
    "note": "SASS mix 538.7 M DFMA + 191.1 M DMUL + 67.9 M DADD warp instructions x 28.90 active lanes = 38.6 GFLOP per launch; "
            "3.86 warps per SM sub-partition (128 registers, four CTAs per SM)",
}
METRIC = "candidate iLQR solves/sec"
UNIT = "solves/s"


def flops_per_iteration(N):
    return 927 * N + 177  # dense algorithmic FP64 flops per iLQR iteration


def bytes_per_solve(N):
    return 48 * N + 132  # algorithmic HBM bytes per candidate solve


def measured_peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        return {}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index=0):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "20"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True,
            )  # fmt: skip
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
                power.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {
            "sm_mhz": float(np.median(sm)) if sm else None,
            "sm_max_mhz": float(max(mx)) if mx else None,
            "power_w_max": float(max(power)) if power else None,
            "samples": len(sm),
            "reasons": sorted(reasons),
        }


def make_workload(B, N, rank):
    from ilqr_iterative_tasks_b200 import _abi, synth

    d = synth.flat_candidates(B, N, seed=synth.SEED + rank)
    return d, _abi.obstacle_array(d["obs"], d["mode"])


def cpu_arm(params, B, N, seconds_target=12.0):
    """The oracle port on all host cores over a bounded sample of the workload."""
    from oracle import oracle

    cores = os.cpu_count() or 1
    probe_n = min(B, 4096 * cores)
    d, obs = make_workload(probe_n, N, 0)
    t0 = time.perf_counter()
    oracle.ilqr_batch(params, d["x0"], d["x_term"], d["lamb_in"], obs, nthreads=cores)
    rate = probe_n / (time.perf_counter() - t0)
    sample = int(min(B, max(probe_n, rate * seconds_target)))
    d, obs = make_workload(sample, N, 0)
    t0 = time.perf_counter()
    _, _, _, iters = oracle.ilqr_batch(params, d["x0"], d["x_term"], d["lamb_in"], obs, nthreads=cores)
    dt = time.perf_counter() - t0
    return {
        "value": sample / dt, "unit": UNIT, "cores": cores, "kind": "port",
        "sample": "first %d of the %d candidates of rank 0's batch, C oracle port (oracle/i2lqr_oracle.c), %d host threads, %.1f s"
        % (sample, B, cores, dt),
        "iters_per_solve": float(iters.mean()), "seconds": dt,
    }  # fmt: skip


def run_reference(args):
    """--impl reference: the CPU implementation of the path on the host cores (rank 0 only)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from ilqr_iterative_tasks_b200 import _abi
    from oracle import oracle

    oracle.build()
    params = _abi.default_params()
    N, B = params.num_horizon, args.batch
    cores = os.cpu_count() or 1
    # each step = a bounded sample of the per-GPU batch so that warmup+steps end within minutes
    d, obs = make_workload(min(B, 8192 * cores), N, 0)
    t0 = time.perf_counter()
    oracle.ilqr_batch(params, d["x0"], d["x_term"], d["lamb_in"], obs, nthreads=cores)
    rate = d["x0"].shape[0] / (time.perf_counter() - t0)
    budget = 120.0 / max(1, args.steps + args.warmup)
    sample = int(min(B, max(4096, rate * min(budget, 15.0))))
    d, obs = make_workload(sample, N, 0)
    for _ in range(args.warmup):
        oracle.ilqr_batch(params, d["x0"], d["x_term"], d["lamb_in"], obs, nthreads=cores)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        oracle.ilqr_batch(params, d["x0"], d["x_term"], d["lamb_in"], obs, nthreads=cores)
    dt = time.perf_counter() - t0
    value = sample * args.steps / dt
    desc = "first %d of the %d candidates per step, C oracle port of the reference's NumPy path, %d host threads" % (sample, B, cores)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(B, N, args.gpus, 1 if args.no_pipeline else max(1, args.pipeline), args.pre_warm_s),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": desc},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "the reference is pure Python (29.7 solves/s/core measured in the build container, BASELINE.md) and cannot "
                "travel to the GPU box; this arm times its C restatement (oracle/), which is ~500x faster per core",
    }  # fmt: skip
    print(json.dumps(line), flush=True)


def workload_config(B, N, gpus, n_pipe=4, pre_warm_s=0.0):
    return {
        "pre_warm": "%.1f s of untimed steps before the W warm-up steps (SM clock at its load state when the timed region starts)" % pre_warm_s,
        "workload": "BASELINE config 5: flat batch of independent candidate iLQR solves, synthetic obstacles near the path",
        "candidates_per_gpu": B, "global_candidates": B * gpus, "num_horizon": N, "state_dim": 4, "input_dim": 2,
        "seed": 20260101, "parallelism": "candidates sharded by rank, no data-path collective; steps pipelined over %d streams per GPU" % n_pipe,
        "l2": "per-step working set (inputs 120 B + outputs 300 B per candidate) is %.0f MB > 126 MB L2" % (B * 420 / 1e6),
        "inputs": "every step reads the SAME resident input batch (outputs rotate over the handles' buffers); the kernel is FP64-bound and "
                  "the 126 MB of inputs are re-read from HBM every step because the 440 MB per-step working set exceeds L2",
    }  # fmt: skip


def latency_probe(device, max_seconds=60.0):
    """p50 / p90 latency of one controller step at the BASELINE config 2 shape (K=64, L=8) via the C ABI."""
    from ilqr_iterative_tasks_b200 import _abi, engine, episodes, synth

    p = _abi.default_params(num_ss_points=64, num_ss_iter=8)
    lap0, _ = synth.initial_lap()
    obs = _abi.obstacle_array(np.array([[31.0, -3.0, 8.0, 6.0, 0.0]]), [0])
    eb = episodes.EpisodeBatch(p, np.zeros((1, 4)), obstacles=obs, initial_lap=lap0, device=device)
    t0 = time.perf_counter()
    laps = [lap0]
    for _ in range(8):  # fill the safe set with 8 learned laps
        n, lx, _ = eb.run_lap(want_traj=True)
        laps.append(lx[0, : int(n[0]) + 1].copy())
        if time.perf_counter() - t0 > max_seconds:
            break
    eb.close()
    eng = engine.Engine(p, device=device, n_episodes=1, lap_capacity=128)
    for lp in laps[-8:]:
        eng.add_lap(0, lp)
    eng.set_obstacle(0, _abi.Obstacle(31.0, -3.0, 8.0, 6.0, 0.0, 0, 1))
    states = laps[-1][:-6]  # the solve steps of the last lap
    times, solves, dev_ms = [], [], []
    for rep in range(3):
        for x in states:
            t1 = time.perf_counter()
            eng.solve_step(x.reshape(1, 4))
            times.append((time.perf_counter() - t1) * 1e3)
            st = eng.last_stats()
            solves.append(st["solves"])
            dev_ms.append(st["kernel_ms"])
    spec = eng.spec_stats()
    eng.close()
    times = np.array(times[len(states):])  # first pass = warm-up
    return {
        "p50_ms": float(np.percentile(times, 50)), "p90_ms": float(np.percentile(times, 90)), "steps_timed": int(times.size),
        "device_p50_ms": float(np.percentile(dev_ms[len(states):], 50)), "launches_per_step": int(st["launches"]),
        "solves_per_step": float(np.mean(solves)), "stored_laps": len(laps[-8:]),
        "speculation": {"hit_rate": spec["hit_rate"], "counters": spec,
                        "note": "one-launch step (spec_step.cuh): candidates served from speculative solves whose lambda_in matched bit for bit"},
        "workload": "BASELINE config 2: one controller step (state in -> u, x_pred, u_pred on host), K=64, L=8, stock scenario",
        "reference_cpu_p50_ms": 9100.0, "reference_note": "Python reference, 1 core, build container (BASELINE.md)",
    }  # fmt: skip


def episodes_probe(device, rank, world, dist, torch, laps=3):
    """BASELINE config 3: 4096 independent closed-loop episodes (random obstacles / start states, seed 20260101,
    K=8, L=2) sharded over the ranks in contiguous blocks (`dist.shard_range`); at world == 1 this GPU runs the
    512-episode share it would own on an 8-GPU box.  Per lap: wall time (max over ranks) and closed-loop candidate
    solves/s; rank 0 re-runs its first 64 episodes for two laps on the CPU oracle as the parity flag."""
    from ilqr_iterative_tasks_b200 import _abi, episodes, synth
    from ilqr_iterative_tasks_b200 import dist as D

    total = 4096
    sc = synth.episode_scenarios(total)
    lo, hi = D.shard_range(total, rank, world) if world > 1 else (0, 512)
    E = hi - lo
    p = _abi.default_params(num_ss_points=8, num_ss_iter=2)
    lap0, _ = synth.initial_lap()
    obs = _abi.obstacle_array(sc["obs"][lo:hi], sc["mode"][lo:hi])
    t0 = time.perf_counter()
    eb = episodes.EpisodeBatch(p, sc["x0"][lo:hi], obstacles=obs, initial_lap=lap0, device=device)
    setup_s = time.perf_counter() - t0
    rows, keep = [], []
    for lap in range(laps):
        if dist is not None:
            dist.barrier()
        t0 = time.perf_counter()
        lens, lx, _ = eb.run_lap(want_traj=True)
        wall = time.perf_counter() - t0
        st = eb.stats()
        status = eb.status()
        keep.append((lens[:64].copy(), lx[:64].copy()))
        t = torch.tensor([wall, float(st["solves"]), float(st["ilqr_iterations"]), float(lens.sum()), float((status == 0).sum()), float(lens.max())],
                         dtype=torch.float64, device=torch.device("cuda", device))  # fmt: skip
        if dist is not None:
            tm = t.clone()
            dist.all_reduce(tm, op=dist.ReduceOp.MAX)
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
            wall_max, steps_max = float(tm[0]), float(tm[5])
        else:
            wall_max, steps_max = wall, float(t[5])
        rows.append({
            "lap": lap + 1, "wall_s": wall_max, "solves": int(t[1]), "ilqr_iterations": int(t[2]), "solves_per_s": float(t[1]) / wall_max,
            "episode_steps": int(t[3]), "steps_max": int(steps_max), "reached_goal": int(t[4]), "launches_rank0": st["launches"],
        })  # fmt: skip
    eb.close()
    out = {
        "workload": "BASELINE config 3: %d closed-loop episodes, K=8, L=2, N=6, seed 20260101, %s" % (
            total if world > 1 else E, "sharded over %d GPUs (%d per GPU)" % (world, E) if world > 1 else "one GPU's share of the 4096 on an 8-GPU box"),
        "episodes_total": total if world > 1 else E, "episodes_per_gpu": E, "setup_s_rank0": setup_s, "laps": rows,
    }  # fmt: skip
    if rank == 0:
        try:
            out["parity"] = _episodes_parity(p, sc, lap0, keep)
        except Exception as err:  # the checker must never cost the line
            out["parity"] = {"error": str(err)[:200]}
    return out


def _episodes_parity(p, sc, lap0, keep, n_eps=64, n_laps=2):
    """Checker leg (CPU oracle): the first 64 episodes, two laps, closed loop."""
    from concurrent.futures import ThreadPoolExecutor

    from tests import helpers as H

    def one(e):
        ref = H.oracle_closed_loop(p, sc["x0"][e], sc["obs"][e], sc["mode"][e], lap0, n_laps)
        ok, err = True, 0.0
        for lap in range(n_laps):
            lens, lx = keep[lap]
            if int(lens[e]) != len(ref[lap + 1]) - 1:
                return False, float("inf")
            err = max(err, float(np.abs(lx[e, : lens[e] + 1] - ref[lap + 1]).max()))
        return ok, err

    t0 = time.perf_counter()
    with ThreadPoolExecutor(max_workers=min(16, os.cpu_count() or 1)) as ex:
        res = list(ex.map(one, range(n_eps)))
    worst = max(r[1] for r in res)
    return {"episodes_checked": n_eps, "laps_checked": n_laps, "lap_lengths_equal": bool(all(r[0] for r in res)), "max_abs_state_err": worst,
            "ok": bool(all(r[0] for r in res) and worst < 1e-8), "checker": "oracle closed loop (tests/helpers.py), %.1f s" % (time.perf_counter() - t0)}  # fmt: skip


def host_ceiling_probe(dev, dist, torch, nbytes, reps=3, h2d_bytes=0):
    """What the host side of PCIe gives every rank when ALL ranks copy at once: pinned D2H (and H2D) of one step's
    result bytes, barrier-aligned, best of `reps`; max-over-ranks time -> aggregate GB/s.  This is the ceiling `e2e`
    is measured against (the single-NUMA VMs of this pool cap aggregate device-to-host traffic well below N links)."""
    n = int(nbytes)
    hbuf = torch.empty(n, dtype=torch.uint8).pin_memory()
    dbuf = torch.empty(n, dtype=torch.uint8, device=dev)
    out = {}
    world = dist.get_world_size() if dist is not None else 1
    for name, (dst, src) in (("d2h", (hbuf, dbuf)), ("h2d", (dbuf, hbuf))):
        best = float("inf")
        for _ in range(reps + 1):
            torch.cuda.synchronize()
            if dist is not None:
                dist.barrier()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            dst.copy_(src, non_blocking=True)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            t = torch.tensor([dt], dtype=torch.float64, device=dev)
            if dist is not None:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            best = min(best, float(t[0]))
        out[name + "_gbs_per_gpu"] = n / best / 1e9
        out[name + "_gbs_aggregate"] = n * world / best / 1e9
    # duplex: the pipeline uploads the next step's inputs WHILE it downloads results, and both directions share the host's
    # memory system -- so the download rate with a concurrent upload of one step's input bytes is the ceiling that applies
    if h2d_bytes:
        m = int(h2d_bytes)
        hin = torch.empty(m, dtype=torch.uint8).pin_memory()
        din = torch.empty(m, dtype=torch.uint8, device=dev)
        up = torch.cuda.Stream(dev)
        best = float("inf")
        for _ in range(reps + 1):
            torch.cuda.synchronize()
            if dist is not None:
                dist.barrier()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            with torch.cuda.stream(up):
                din.copy_(hin, non_blocking=True)
            hbuf.copy_(dbuf, non_blocking=True)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            t = torch.tensor([dt], dtype=torch.float64, device=dev)
            if dist is not None:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            best = min(best, float(t[0]))
        out["d2h_gbs_aggregate_duplex"] = n * world / best / 1e9
        out["duplex_h2d_bytes"] = m
    out["bytes"] = n
    return out



def bind_to_gpu_numa_node(index, verbose=False):
    """Pins this process to the CPUs the driver reports as local to GPU `index` (NVML's ideal CPU affinity) BEFORE any
    pinned buffer is allocated, so that the end-to-end arm's host buffers live on the GPU's own NUMA node: with
    several ranks per box the host side of PCIe is the bottleneck of `e2e`, and remote-node buffers halve it."""
    if os.environ.get("I2LQR_NO_AFFINITY"):
        return
    try:
        import pynvml

        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        pynvml.nvmlDeviceSetCpuAffinity(h)
        if verbose:
            print("[bench] rank 0 bound to CPUs %s (NVML ideal affinity of GPU %d)" % (sorted(os.sched_getaffinity(0)), index), file=sys.stderr)
    except Exception as err:  # affinity is an optimisation, never a requirement
        if verbose:
            print("[bench] CPU affinity not set: %s" % str(err)[:120], file=sys.stderr)


def run_ours(args):
    import torch

    from ilqr_iterative_tasks_b200 import _abi, engine

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device (there is no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local)
    bind_to_gpu_numa_node(local, verbose=(rank == 0))
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    params = _abi.default_params()
    N, B = params.num_horizon, args.batch
    d, obs = make_workload(B, N, rank)
    # several handles = several streams: consecutive steps are independent batches, so the next steps start while the
    # iteration-count tail of step k (its resume passes: a few 150-iteration stragglers, 1.5 ms of sequential trips) is
    # still draining.  Measured on one B200 at 24 steps: 2 streams 367, 3: 390, 4: 402, 8: 412 M solves/s
    n_pipe = 1 if args.no_pipeline else max(1, args.pipeline)
    # the end-to-end arm keeps at most four caller threads (one handle and one set of pinned buffers each): it is bound by the
    # download direction of PCIe from two threads on, and with eight ranks on a 32-CPU box more threads only oversubscribe the host
    n_host = min(n_pipe, 4)
    engs = [engine.Engine(params, device=local) for _ in range(n_pipe)]
    eng = engs[0]
    dev = torch.device("cuda", local)

    def pinned(a):
        return torch.from_numpy(np.ascontiguousarray(a)).pin_memory()

    h_x0, h_xt, h_lamb = pinned(d["x0"]), pinned(d["x_term"]), pinned(d["lamb_in"])
    h_obs = torch.from_numpy(obs.view(np.uint8).reshape(B, -1).copy()).pin_memory()
    h_outs = [
        {
            "uvar": torch.empty((B, 2, N), dtype=torch.float64).pin_memory(),
            "xvar": torch.empty((B, 4, N + 1), dtype=torch.float64).pin_memory(),
            "lamb_out": torch.empty(B, dtype=torch.float64).pin_memory(),
            "iters": torch.empty(B, dtype=torch.int32).pin_memory(),
        }
        for _ in range(n_host)
    ]
    # ---- HBM-resident arm (value + roofline)
    g_x0, g_xt, g_lamb, g_obs = h_x0.to(dev), h_xt.to(dev), h_lamb.to(dev), h_obs.to(dev)
    g_out = [
        (
            torch.empty((B, 2, N), dtype=torch.float64, device=dev),
            torch.empty((B, 4, N + 1), dtype=torch.float64, device=dev),
            torch.empty(B, dtype=torch.float64, device=dev),
            torch.empty(B, dtype=torch.int32, device=dev),
        )
        for _ in range(n_pipe)
    ]
    streams = [torch.cuda.Stream(device=dev) for _ in range(n_pipe)]

    def device_step(k):
        e, (g_u, g_x, g_l, g_i), st = engs[k % n_pipe], g_out[k % n_pipe], streams[k % n_pipe]
        e.solve_batch_device(
            B, g_x0.data_ptr(), g_xt.data_ptr(), g_lamb.data_ptr(), g_obs.data_ptr(), B,
            g_u.data_ptr(), g_x.data_ptr(), g_l.data_ptr(), g_i.data_ptr(), stream=st.cuda_stream,
        )  # fmt: skip

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    fp64_peak = engine.measure_fp64_peak(local, reps=5)
    # clocks / throttle reasons are sampled from the warm-up on (nvidia-smi needs ~100 ms to deliver its first sample and the timed
    # region of the default run is 27 ms), through the timed region and the >= 2 s `sustained` repetition of the same loop
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    # W warm-up steps, and at least one on every handle: a handle allocates its device buffers on first use, which must
    # not happen inside the timed region (it did once, with 4 handles and W = 3: 31 ms per step instead of 2.6)
    # (before them, `--pre-warm-s` seconds of the same untimed steps: on a box whose GPU has been idle -- the driver runs the
    #  CPU-only reference arm first -- the first ~100 ms of load run below the load clock, and a 50 ms timed region that
    #  starts 30 ms after the first launch measured 6 % low once, 400 instead of 426 M solves/s)
    t_ramp = time.perf_counter()
    k_ramp = 0
    while time.perf_counter() - t_ramp < args.pre_warm_s:
        device_step(k_ramp)
        k_ramp += 1
        if k_ramp % n_pipe == 0:
            streams[0].synchronize()
    for k in range(max(args.warmup, n_pipe)):
        device_step(k)
    barrier()
    # isolated launches (one stream, back to back): the per-launch duration ncu also sees
    iso = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    iso[0].record(streams[0])
    for k in range(3):
        device_step(0)
    iso[1].record(streams[0])
    streams[0].synchronize()
    isolated_ms = iso[0].elapsed_time(iso[1]) / 3
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(streams[0])
    for st in streams[1:]:
        st.wait_event(ev0)
    for k in range(args.steps):
        device_step(k)
    for st in streams[1:]:
        join = torch.cuda.Event()
        join.record(st)
        streams[0].wait_event(join)
    ev1.record(streams[0])
    streams[0].synchronize()
    barrier()
    total_ms = ev0.elapsed_time(ev1)
    step_ms = [total_ms / args.steps]
    st = eng.sync_stats(streams[0].cuda_stream)
    # ---- sustained: the same rotation for >= args.sustain_s seconds (hundreds of steps), clocks / power sampled alongside
    sustained = None
    if args.sustain_s > 0:
        n_sus = max(args.steps, int(args.sustain_s * 1e3 / max(step_ms[0], 1e-3)) + 1)
        ev0.record(streams[0])
        for st_ in streams[1:]:
            st_.wait_event(ev0)
        for k in range(n_sus):
            device_step(k)
        for st_ in streams[1:]:
            join = torch.cuda.Event()
            join.record(st_)
            streams[0].wait_event(join)
        ev1.record(streams[0])
        streams[0].synchronize()
        barrier()
        sus_ms = ev0.elapsed_time(ev1)
        tsus = torch.tensor([sus_ms], dtype=torch.float64, device=dev)
        if dist is not None:
            dist.all_reduce(tsus, op=dist.ReduceOp.MAX)
        sustained = {"steps": n_sus, "seconds": float(tsus[0]) * 1e-3, "ms_per_step": float(tsus[0]) / n_sus,
                     "value": B * world * n_sus / (float(tsus[0]) * 1e-3), "unit": UNIT}  # fmt: skip
    clocks = sampler.stop() if rank == 0 else None
    if clocks is not None:
        clocks["window"] = "warm-up + timed region + sustained repetition of the same loop"
    iters_total = st["ilqr_iterations"]  # per step (identical every step)
    kernel_ms = isolated_ms
    launches_per_step = st["launches"]
    # ---- end-to-end arm: host buffers through the C-ABI host call, one host thread per handle
    import threading

    x0_np, xt_np, lamb_np = h_x0.numpy(), h_xt.numpy(), h_lamb.numpy()
    obs_np = h_obs.numpy().view(_abi.OBSTACLE_DTYPE).reshape(B)  # the pinned copy: a pageable upload would serialise the pipeline
    outs_np = [{k: v.numpy() for k, v in ho.items()} for ho in h_outs]

    def host_steps(i, n):
        for _ in range(n):
            engs[i].solve_batch(x0_np, xt_np, lamb_np, obs_np, out=outs_np[i])

    for i in range(n_host):
        host_steps(i, 1)
    barrier()
    share = [args.steps // n_host + (1 if i < args.steps % n_host else 0) for i in range(n_host)]
    threads = [threading.Thread(target=host_steps, args=(i, share[i])) for i in range(n_host)]
    t0 = time.perf_counter()
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    out_np = outs_np[0]
    parity_ok = bool(np.array_equal(out_np["iters"], g_out[0][3].cpu().numpy()) and np.array_equal(out_np["xvar"], g_out[0][1].cpu().numpy()))
    # ---- e2e_min: same call with the trajectory outputs declined (NULL uvar / xvar): per candidate only lamb_out and the
    # iteration count come back (12 B instead of 300 B), which is what a caller needs to rank candidates
    min_np = [{"lamb_out": o["lamb_out"], "iters": o["iters"]} for o in outs_np]

    def host_steps_min(i, n):
        for _ in range(n):
            engs[i].solve_batch(x0_np, xt_np, lamb_np, obs_np, want=("lamb_out", "iters"), out=min_np[i])

    for i in range(n_host):
        host_steps_min(i, 1)
    barrier()
    threads = [threading.Thread(target=host_steps_min, args=(i, share[i])) for i in range(n_host)]
    t0 = time.perf_counter()
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    torch.cuda.synchronize()
    e2e_min_s = time.perf_counter() - t0
    barrier()
    ceiling = host_ceiling_probe(dev, dist, torch, B * (16 * N + 32 * (N + 1) + 8 + 4), h2d_bytes=B * 120)
    # ---- max over ranks
    t = torch.tensor([total_ms, e2e_s * 1e3, kernel_ms, float(iters_total), e2e_min_s * 1e3], dtype=torch.float64, device=dev)
    if dist is not None:
        tmax = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone()
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        total_ms, e2e_ms, e2e_min_ms = float(tmax[0]), float(tmax[1]), float(tmax[4])
    else:
        e2e_ms, e2e_min_ms = e2e_s * 1e3, e2e_min_s * 1e3
    # ---- the other BASELINE configs that need a driver-run record (all ranks take part)
    episodes = None
    if not args.no_episodes:
        try:
            episodes = episodes_probe(local, rank, world, dist, torch)
        except Exception as err:
            episodes = {"error": str(err)[:300]}
    split = None
    if world >= 2 and not args.no_split:
        try:
            from tests import multi_gpu_split_step

            split = multi_gpu_split_step.run_check(rank, world, local, max_parity_steps=24)
        except Exception as err:
            split = {"error": str(err)[:300], "bit_identical": False}
    if rank == 0:
        peaks = measured_peaks()
        hbm_peak = peaks.get("hbm_gbs")
        value = B * world * args.steps / (total_ms * 1e-3)
        e2e_value = B * world * args.steps / (e2e_ms * 1e-3)
        flops = iters_total * flops_per_iteration(N)
        avg_ms = float(t[0]) / args.steps  # this rank's average launch duration over the timed region
        achieved = flops / (avg_ms * 1e-3) / 1e12
        h2d = B * (32 + 32 + 8 + 48)
        d2h = B * (16 * N + 32 * (N + 1) + 8 + 4)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": workload_config(B, N, world, n_pipe, args.pre_warm_s),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": e2e_ms / args.steps,
                    "api": "i2lqr_solve_batch (ctypes), pinned host buffers, %d host threads each driving its own handle" % n_host,
                    "host_ceiling": ceiling,
                    "d2h_gbs_aggregate": d2h * world / (e2e_ms / args.steps * 1e-3) / 1e9,
                    "frac_of_d2h_ceiling": d2h * world / (e2e_ms / args.steps * 1e-3) / 1e9 / ceiling["d2h_gbs_aggregate"],
                    "frac_of_duplex_ceiling": d2h * world / (e2e_ms / args.steps * 1e-3) / 1e9 / ceiling["d2h_gbs_aggregate_duplex"],
                    "ceiling_note": "host_ceiling = all ranks copying one step's result bytes device-to-host at once (pinned, barrier-aligned, "
                                    "max over ranks); e2e is download-bound, so d2h_gbs_aggregate / host_ceiling.d2h_gbs_aggregate is how close the "
                                    "pipeline runs to what the box gives; d2h_gbs_aggregate_duplex = the same download with one step's input "
                                    "bytes going up at the same time, as they do in the pipeline"},
            "e2e_min": {"value": B * world * args.steps / (e2e_min_ms * 1e-3), "unit": UNIT, "ms_per_step": e2e_min_ms / args.steps,
                        "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": B * 12,
                        "note": "VARIANT, not the headline: same host call with uvar / xvar declined (NULL) -- only lamb_out and the iteration "
                                "count per candidate come back"},
            "gpu_launches": int(launches_per_step * args.steps),
            "clocks": clocks,
            "roofline": {
                "bound": "fp64_fma", "achieved": achieved, "peak": fp64_peak, "unit": "TFLOP/s", "frac": achieved / fp64_peak,
                "traffic": NCU_TRAFFIC_BYTES if (B == 1 << 20 and N == 6) else None,
                "traffic_note": "dram__bytes_read.sum + dram__bytes_write.sum of the pass-0 launch (2^20 candidates), one ncu --set full "
                                "capture: profiles/r02_solve_kernel_v8_lean_pass0_summary.txt; algorithmic bytes per launch = %d" % (B * bytes_per_solve(N)),
                "kernel": "solve_kernel<double, kQR=false, kScratch=6 (lean tensor-memory placement: gains and terminal target in tensor memory, 128 registers, four CTAs per SM), kMode=0, 128 threads>", "kernel_ms": avg_ms,
                "isolated": {"launch_ms": kernel_ms, "achieved": flops / (kernel_ms * 1e-3) / 1e12,
                             "frac": flops / (kernel_ms * 1e-3) / 1e12 / fp64_peak,
                             "note": "one launch at a time on one stream (what ncu sees); the timed region pipelines "
                                     "consecutive steps over %d streams so one step's straggler tail overlaps the next step" % n_pipe},
                "peak_nominal": 148 * 64 * 2 * 1.965e-3, "frac_of_nominal": achieved / (148 * 64 * 2 * 1.965e-3),
                "peak_nominal_note": "148 SMs x 64 FP64 FMA lanes x 2 flops x 1.965 GHz",
                "executed": NCU_EXECUTED,
                "ilqr_iterations_per_step": int(iters_total), "iters_per_solve": iters_total / B,
                "flops_per_iteration": flops_per_iteration(N),
                "peak_source": "dependent-free DFMA micro-benchmark (i2lqr_measure_fp64_peak) on this GPU, same run; "
                               "MEASURED_PEAKS.json has no FP64 figure (nominal B200 FP64 ~37 TFLOP/s)",
                "hbm": {"achieved_gbs": B * bytes_per_solve(N) / (avg_ms * 1e-3) / 1e9, "peak_gbs": hbm_peak,
                        "bytes_per_solve": bytes_per_solve(N), "peak_source": "MEASURED_PEAKS.json"},
            },
            "step_ms": step_ms, "e2e_matches_device_arm": parity_ok,
            "sustained": sustained, "episodes": episodes, "split_step": split,
        }  # fmt: skip
        if world == 1 and not args.no_cpu:
            line["cpu_baseline"] = cpu_arm(params, B, N)
        if world == 1 and not args.no_latency:
            try:
                line["latency"] = latency_probe(local)
            except Exception as err:  # the probe must never cost the headline number
                line["latency"] = {"error": str(err)[:200]}
        print(json.dumps(line), flush=True)
    for e in engs:
        e.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=1 << 20, help="candidates per GPU per step")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-latency", action="store_true", help="skip the controller-step latency probe")
    ap.add_argument("--no-pipeline", action="store_true", help="one handle / one stream: steps strictly back to back")
    ap.add_argument("--pipeline", type=int, default=6, help="handles (streams) that consecutive steps rotate over")
    ap.add_argument("--pre-warm-s", type=float, default=0.4, help="seconds of untimed steps before the W warm-up steps (SM clock ramp)")
    ap.add_argument("--sustain-s", type=float, default=2.0, help="seconds of the `sustained` leg (0 = skip)")
    ap.add_argument("--no-episodes", action="store_true", help="skip the BASELINE config 3 leg (`episodes` object)")
    ap.add_argument("--no-split", action="store_true", help="skip the NCCL split-step leg (`split_step` object, world >= 2)")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
