"""Measurements of the BASELINE.json configs that are not the bench.py headline (config 5):

    python bench_configs.py stock       # config 1: the reference's own 10-lap run through the drop-in classes
    python bench_configs.py latency     # config 2: one closed-loop lap, K=64, L=8: per-step latency on 1 GPU
    python bench_configs.py episodes    # config 3: independent closed-loop episodes (512 per GPU), K=8, L=2
    python bench_configs.py horizon     # config 4a: horizon sweep N=10..200, flat batch with obstacle barrier

Each prints one JSON line; the committed copies live under profiles/.  All calls go through the C ABI.
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from ilqr_iterative_tasks_b200 import _abi, engine, episodes, synth  # noqa: E402


def flops_per_iteration(N):
    return 927 * N + 177


# measured DRAM traffic of the streaming placement per candidate, stage and iLQR iteration (doubles): dram__bytes of the ncu capture
# of the N = 50 launch / (executed iterations x 50 x 8); the count from the code is 6 + 10 + 16 + 6 + 6..12 = 44..50
STREAM_DOUBLES_PER_STAGE = 56.0


def stock(device=0, laps=10, K=8, L=2):
    """BASELINE config 1: the reference's own run -- `ilqr_test.py --lap-number 10 --num-ss-iters 2
    --num-ss-points 8` -- driven through the drop-in classes exactly as that script drives the reference
    (one `calc_input` = one controller step on the GPU; plant and lap bookkeeping in the host loop)."""
    import ilqr_iterative_tasks_b200 as pkg

    pkg.install_dropin()
    from utils import base  # the drop-in

    dt, N = 1, 6
    ego = base.KineticBicycle(False, system_param=base.KineticBicycleParam())
    ego.set_state([0, 0, 0, 0])
    ego.set_timestep(dt)
    ego.get_traj()
    ego.set_zero_noise()
    ctrl = base.iLqr(base.iLqrParam(num_ss_points=K, num_ss_iter=L, timestep=dt, num_horizon=N), obstacle=base.Obstacle(31, -3, 8, 6),
                     system_param=base.KineticBicycleParam())
    ctrl.add_trajectory(ego.xcl, ego.ucl)
    ctrl.set_timestep(dt)
    ego.set_ctrl_policy(ctrl)
    sim = base.Simulator()
    sim.set_robotic(ego)
    sim.set_timestep(dt)
    sim.set_traj()
    import contextlib
    import io

    t0 = time.perf_counter()
    with contextlib.redirect_stdout(io.StringIO()):  # Simulator.sim prints "iteration: k finished" like the reference
        for lap in range(laps):
            sim.sim(lap, sim_time=50)
            ego.data["state"][-1][-1, :] = ego.xcl[-1, :]
            ctrl.add_trajectory(ego.data["state"][-1], ego.data["input"][-1])
    wall = time.perf_counter() - t0
    solver = np.concatenate([np.asarray(d).ravel() for d in ego.diagnostics["solver_time"]])
    return {
        "config": "BASELINE config 1: ilqr_test.py --lap-number %d --num-ss-iters %d --num-ss-points %d through the drop-in" % (laps, L, K),
        "wall_s": wall, "lap_steps": [int(len(s)) - 1 for s in ego.data["state"]],
        "controller_step_ms_p50": float(np.median(solver) * 1e3), "controller_steps": int(solver.size),
        "reference": {"wall_s": 302.0, "lap_steps": [54, 29, 23, 23, 22, 22, 22, 22, 22, 22],
                      "note": "unmodified Python reference in the build container, one core (SURVEY.md 8c)"},
    }  # fmt: skip


def latency(device=0):
    import bench

    out = bench.latency_probe(device, max_seconds=120.0)
    out["config"] = "BASELINE config 2"
    return out


def episodes_cfg(device=0, E=512, laps=3):
    sc = synth.episode_scenarios(4096)
    rank = int(os.environ.get("RANK", "0"))
    lo = (rank * E) % 4096 if E < 4096 else 0
    p = _abi.default_params(num_ss_points=8, num_ss_iter=2)
    lap0, _ = synth.initial_lap()
    obs = _abi.obstacle_array(sc["obs"][lo : lo + E], sc["mode"][lo : lo + E])
    t0 = time.perf_counter()
    eb = episodes.EpisodeBatch(p, sc["x0"][lo : lo + E], obstacles=obs, initial_lap=lap0, device=device)
    setup_s = time.perf_counter() - t0
    rows = []
    for lap in range(laps):
        t0 = time.perf_counter()
        lens = eb.run_lap()
        wall = time.perf_counter() - t0
        st = eb.stats()
        status = eb.status()
        rows.append({
            "lap": lap + 1, "wall_s": wall, "kernel_ms": st["kernel_ms"], "launches": st["launches"],
            "solves": st["solves"], "ilqr_iterations": st["ilqr_iterations"],
            "steps_mean": float(lens.mean()), "steps_min": int(lens.min()), "steps_max": int(lens.max()),
            "reached_goal": int((status == 0).sum()), "solves_per_s": st["solves"] / wall,
            "episode_steps_per_s": float(lens.sum()) / wall,
        })  # fmt: skip
    eb.close()
    return {
        "config": "BASELINE config 3 (this GPU's share: %d of 4096 episodes)" % E, "episodes": E, "K": 8, "L": 2, "N": 6,
        "setup_s": setup_s, "laps": rows,
        "reference_cpu": "Python reference: ~1.4 s per controller step per episode on one core (BASELINE.md)",
    }


def horizon_sweep(device=0):
    """BASELINE config 4a.  Per horizon: (1) the kernel on device-resident batches large enough that the 150-iteration
    stragglers do not dominate (CUDA events; algorithmic FP64 fraction AND, for the streaming placement of N > 25, the
    scratch traffic against the measured HBM copy bandwidth); (2) the same solve through the host-pointer call with a
    65536-candidate batch (copies inside); (3) the optional FP32 build's tolerance gate."""
    import torch

    rows = []
    peak = engine.measure_fp64_peak(device)
    try:
        hbm_peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        hbm_peak = None
    dev = torch.device("cuda", device)
    for N, B in ((6, 1 << 20), (10, 1 << 20), (25, 1 << 20), (50, 1 << 20), (100, 1 << 20), (200, 1 << 20)):
        d = synth.flat_candidates(B, N, seed=synth.SEED + N)
        p = _abi.default_params(num_horizon=N)
        obs = _abi.obstacle_array(d["obs"], d["mode"])
        e = engine.Engine(p, device=device)
        g = {k: torch.from_numpy(np.ascontiguousarray(d[k])).to(dev) for k in ("x0", "x_term", "lamb_in")}
        g_obs = torch.from_numpy(obs.view(np.uint8).reshape(B, -1).copy()).to(dev)
        out = (torch.empty((B, 2, N), dtype=torch.float64, device=dev), torch.empty((B, 4, N + 1), dtype=torch.float64, device=dev),
               torch.empty(B, dtype=torch.float64, device=dev), torch.empty(B, dtype=torch.int32, device=dev))  # fmt: skip
        stream = torch.cuda.Stream(dev)
        best = None
        for rep in range(3):
            e.solve_batch_device(B, g["x0"].data_ptr(), g["x_term"].data_ptr(), g["lamb_in"].data_ptr(), g_obs.data_ptr(), B,
                                 out[0].data_ptr(), out[1].data_ptr(), out[2].data_ptr(), out[3].data_ptr(), stream=stream.cuda_stream)
            st = e.sync_stats(stream.cuda_stream)
            if rep and (best is None or st["kernel_ms"] < best["kernel_ms"]):
                best = st
        tf = best["ilqr_iterations"] * flops_per_iteration(N) / (best["kernel_ms"] * 1e-3) / 1e12
        finite = bool(torch.isfinite(out[1]).all())
        row = {
            "N": N, "device_resident": {
                "B": B, "kernel_ms": best["kernel_ms"], "launches": best["launches"], "solves_per_s": B / (best["kernel_ms"] * 1e-3),
                "iters_per_solve": best["ilqr_iterations"] / B, "algorithmic_tflops": tf, "frac_of_fp64_peak": tf / peak, "finite": finite,
            },
        }  # fmt: skip
        if N > 25:  # streaming placement: DRAM traffic per stage and iteration from the ncu capture (profiles/r02_stream_n50_summary.txt)
            gbs = best["ilqr_iterations"] * N * STREAM_DOUBLES_PER_STAGE * 8 / (best["kernel_ms"] * 1e-3) / 1e9
            row["device_resident"].update({"scratch_traffic_gbs": gbs, "frac_of_hbm_peak": gbs / hbm_peak if hbm_peak else None,
                                           "placement": "global memory streamed through a shared-memory ring with cp.async.bulk"})  # fmt: skip
        else:
            row["device_resident"]["placement"] = "trajectory in shared memory, gains in tensor memory"
        del g, g_obs, out
        torch.cuda.empty_cache()
        if N > 25:
            # ---- the same kernel on a 2^22-candidate batch.  A streaming launch ends with the 0.09 % of candidates that run all 150
            # iterations -- 150 x N stages x 1.9 us of sequential work each (14 ms at N = 50) after the last fresh candidate was
            # taken, i.e. a quarter of a 2^20-candidate launch at a few lanes per warp; four times the candidates quarter that share
            # and show the rate the bulk of the launch runs at.  Both rows are reported.
            BL = 1 << 22
            dl = synth.flat_candidates(BL, N, seed=synth.SEED + N)
            obs_l = _abi.obstacle_array(dl["obs"], dl["mode"])
            gl = {k: torch.from_numpy(np.ascontiguousarray(dl[k])).to(dev) for k in ("x0", "x_term", "lamb_in")}
            gl_obs = torch.from_numpy(obs_l.view(np.uint8).reshape(BL, -1).copy()).to(dev)
            out_l = (torch.empty((BL, 2, N), dtype=torch.float64, device=dev), torch.empty((BL, 4, N + 1), dtype=torch.float64, device=dev),
                     torch.empty(BL, dtype=torch.float64, device=dev), torch.empty(BL, dtype=torch.int32, device=dev))  # fmt: skip
            best_l = None
            for rep in range(2):
                e.solve_batch_device(BL, gl["x0"].data_ptr(), gl["x_term"].data_ptr(), gl["lamb_in"].data_ptr(), gl_obs.data_ptr(), BL,
                                     out_l[0].data_ptr(), out_l[1].data_ptr(), out_l[2].data_ptr(), out_l[3].data_ptr(), stream=stream.cuda_stream)
                st = e.sync_stats(stream.cuda_stream)
                if rep and (best_l is None or st["kernel_ms"] < best_l["kernel_ms"]):
                    best_l = st
            tfl = best_l["ilqr_iterations"] * flops_per_iteration(N) / (best_l["kernel_ms"] * 1e-3) / 1e12
            gbl = best_l["ilqr_iterations"] * N * STREAM_DOUBLES_PER_STAGE * 8 / (best_l["kernel_ms"] * 1e-3) / 1e9
            row["device_resident_large"] = {
                "B": BL, "kernel_ms": best_l["kernel_ms"], "launches": best_l["launches"], "solves_per_s": BL / (best_l["kernel_ms"] * 1e-3),
                "iters_per_solve": best_l["ilqr_iterations"] / BL, "algorithmic_tflops": tfl, "frac_of_fp64_peak": tfl / peak,
                "scratch_traffic_gbs": gbl, "frac_of_hbm_peak": gbl / hbm_peak if hbm_peak else None,
                "finite": bool(torch.isfinite(out_l[1]).all()),
            }  # fmt: skip
            del gl, gl_obs, out_l, dl, obs_l
            torch.cuda.empty_cache()
        # ---- host-pointer call, 65536 candidates, pinned buffers (pageable memory would serialise and dominate the timing)
        Bh = 1 << 16

        def pin(a):
            return torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()

        dh = {k: pin(v[:Bh]) for k, v in d.items()}
        obs_h = np.frombuffer(pin(obs[:Bh].view(np.uint8)), dtype=_abi.OBSTACLE_DTYPE)
        bufs = {"uvar": pin(np.empty((Bh, 2, N))), "xvar": pin(np.empty((Bh, 4, N + 1))), "lamb_out": pin(np.empty(Bh)), "iters": pin(np.empty(Bh, dtype=np.int32))}
        e.solve_batch(dh["x0"], dh["x_term"], dh["lamb_in"], obs_h, out=bufs)
        t0 = time.perf_counter()
        o64 = e.solve_batch(dh["x0"], dh["x_term"], dh["lamb_in"], obs_h, out=bufs)
        wall = time.perf_counter() - t0
        o64 = {k: v.copy() for k, v in o64.items()}
        e.close()
        row["host_call"] = {"B": Bh, "wall_ms": wall * 1e3, "solves_per_s": Bh / wall,
                            "bytes_moved": int(Bh * (bytes_in_per_solve() + 48 * N + 44))}  # fmt: skip
        # ---- optional FP32 build: tolerance gate against the FP64 result of the same inputs
        e32 = engine.Engine(_abi.default_params(num_horizon=N, precision=1), device=device)
        o32 = e32.solve_batch(dh["x0"], dh["x_term"], dh["lamb_in"], obs_h)
        t0 = time.perf_counter()
        e32.solve_batch(dh["x0"], dh["x_term"], dh["lamb_in"], obs_h, out=bufs)
        wall32 = time.perf_counter() - t0
        e32.close()
        with np.errstate(divide="ignore"):
            same = (o32["iters"] == o64["iters"]) & (np.round(np.log10(o32["lamb_out"])) == np.round(np.log10(o64["lamb_out"])))
        ax = tuple(range(1, 3))
        err = np.maximum(
            np.abs(o32["xvar"] - o64["xvar"]).max(axis=ax) / np.maximum(1, np.abs(o64["xvar"]).max(axis=ax)),
            np.abs(o32["uvar"] - o64["uvar"]).max(axis=ax) / np.maximum(1, np.abs(o64["uvar"]).max(axis=ax)),
        )
        row["fp32"] = {"wall_ms": wall32 * 1e3, "same_discrete_path": float(same.mean()),
                       "within_1e-4_of_fp64_where_path_matches": float((err[same] <= 1e-4).mean()), "median_rel_err": float(np.median(err[same]))}  # fmt: skip
        rows.append(row)
    return {"config": "BASELINE config 4a: horizon sweep of the flat candidate solve (obstacle barrier on), FP64 vs FP32 gate",
            "fp64_peak_tflops": peak, "hbm_peak_gbs": hbm_peak, "rows": rows}


def bytes_in_per_solve():
    return 32 + 32 + 8 + 48


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "latency"
    dev = int(os.environ.get("LOCAL_RANK", "0"))
    fn = {"stock": stock, "latency": latency, "episodes": episodes_cfg, "horizon": horizon_sweep}[what]
    if what == "episodes" and len(sys.argv) > 2:  # episodes on this GPU (default: its 512-episode share of the 4096)
        print(json.dumps(episodes_cfg(dev, E=int(sys.argv[2]))))
    else:
        print(json.dumps(fn(dev)))
