"""Developer probe: device-resident throughput of the flat solve at one horizon for several batch sizes and scratch placements.
usage: python tools/horizon_probe.py N B [reps]   (environment: I2LQR_SCRATCH, I2LQR_STREAM_BLOCKS, I2LQR_ITER_CAP ...)"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ilqr_iterative_tasks_b200 import _abi, engine, synth  # noqa: E402

N, B = int(sys.argv[1]), int(sys.argv[2])
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
d = synth.flat_candidates(B, N, seed=synth.SEED + N)
p = _abi.default_params(num_horizon=N)
obs = _abi.obstacle_array(d["obs"], d["mode"])
dev = torch.device("cuda", 0)
g = {k: torch.from_numpy(np.ascontiguousarray(d[k])).to(dev) for k in ("x0", "x_term", "lamb_in")}
g_obs = torch.from_numpy(obs.view(np.uint8).reshape(B, -1).copy()).to(dev)
u = torch.empty((B, 2, N), dtype=torch.float64, device=dev)
x = torch.empty((B, 4, N + 1), dtype=torch.float64, device=dev)
lam = torch.empty(B, dtype=torch.float64, device=dev)
it = torch.empty(B, dtype=torch.int32, device=dev)
e = engine.Engine(p)
peak = engine.measure_fp64_peak(0)
st = torch.cuda.Stream(dev)
best = None
for r in range(reps + 1):
    e.solve_batch_device(B, g["x0"].data_ptr(), g["x_term"].data_ptr(), g["lamb_in"].data_ptr(), g_obs.data_ptr(), B, u.data_ptr(), x.data_ptr(),
                         lam.data_ptr(), it.data_ptr(), stream=st.cuda_stream)
    s = e.sync_stats(st.cuda_stream)
    if r and (best is None or s["kernel_ms"] < best["kernel_ms"]):
        best = s
tf = best["ilqr_iterations"] * (927 * N + 177) / (best["kernel_ms"] * 1e-3) / 1e12
bytes_iter = best["ilqr_iterations"] * 57 * N * 8
print("N %d B %d: %.2f ms, %.2f M solves/s, iters/solve %.1f (max %d), %.2f TFLOP/s algorithmic = %.3f of FP64 peak, ~%.0f GB/s of scratch traffic, launches %d"
      % (N, B, best["kernel_ms"], B / best["kernel_ms"] / 1e3, best["ilqr_iterations"] / B, int(it.max()), tf, tf / peak, bytes_iter / (best["kernel_ms"] * 1e-3) / 1e9, best["launches"]))
e.close()
