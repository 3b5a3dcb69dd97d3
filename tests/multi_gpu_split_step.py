"""Launched by tests/test_gpu_multi.py (and by hand: `python -m torch.distributed.run --nproc-per-node 2
--master-addr 127.0.0.1 --master-port 29741 tests/multi_gpu_split_step.py`).

One process per GPU.  Every rank replays recorded controller steps of the reference's K=64 / L=3 run twice:
whole on its own GPU (i2lqr_solve_step) and split over the ranks (i2lqr_split_*, lap rows exchanged with
ncclAllGather), and checks that the two agree bit for bit, outputs and per-candidate trace.  Rank 0 prints
one JSON line with the p50 latency of both forms.
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from ilqr_iterative_tasks_b200 import dist as D  # noqa: E402
from ilqr_iterative_tasks_b200 import engine  # noqa: E402
from tests import helpers as H  # noqa: E402


def run_check(rank, world, local, max_parity_steps=None):
    """The check itself, inside an initialised NCCL process group (bench.py calls it for the `split_step` object of
    its JSON line).  Returns the result dict on every rank; raises AssertionError on any bit difference."""
    g = H.load("steps_k64.npz")
    p = H.step_params(g)
    whole = engine.Engine(p, device=local, n_episodes=1, lap_capacity=128)
    part = engine.Engine(p, device=local, n_episodes=1, lap_capacity=128)
    split = D.SplitStep(part, rank, world, n_episodes=1)
    added, t_whole, t_split, checked = 0, [], [], 0
    for s in list(g["solve_ids"])[:max_parity_steps]:
        n_laps = int(g["step/n_laps"][s])
        while added < n_laps:
            whole.add_lap(0, g["lap%d" % added])
            part.add_lap(0, g["lap%d" % added])
            added += 1
        whole.set_obstacle(0, H.step_obstacle(g, s))
        part.set_obstacle(0, H.step_obstacle(g, s))
        x = g["step/x"][s].reshape(1, 4)
        t0 = time.perf_counter()
        ref = whole.solve_step(x)
        t1 = time.perf_counter()
        out = split.solve_step(x)
        t2 = time.perf_counter()
        ref_tr, tr = whole.step_trace(1), part.step_trace(1)
        for key in ref:
            assert np.array_equal(out[key], ref[key], equal_nan=True), (rank, int(s), key)
        for key in ref_tr:
            assert np.array_equal(tr[key], ref_tr[key], equal_nan=True), (rank, int(s), key)
        checked += 1
    # latency of both forms at the BASELINE config-2 shape (K=64, L=8, eight stored laps): the recorded laps are
    # stored again until eight are active, then recorded states are replayed
    p8 = H.step_params(g)
    p8.num_ss_iter = 8
    whole8 = engine.Engine(p8, device=local, n_episodes=1, lap_capacity=128)
    part8 = engine.Engine(p8, device=local, n_episodes=1, lap_capacity=128)
    split8 = D.SplitStep(part8, rank, world, n_episodes=1)
    for l in range(8):
        lap = g["lap%d" % min(l, added - 1)]
        whole8.add_lap(0, lap)
        part8.add_lap(0, lap)
    for s in list(g["solve_ids"])[:24]:
        whole8.set_obstacle(0, H.step_obstacle(g, s))
        part8.set_obstacle(0, H.step_obstacle(g, s))
        x = g["step/x"][s].reshape(1, 4)
        dist.barrier()
        t0 = time.perf_counter()
        ref = whole8.solve_step(x)
        t1 = time.perf_counter()
        dist.barrier()
        t2 = time.perf_counter()
        out = split8.solve_step(x)
        t3 = time.perf_counter()
        for key in ref:
            assert np.array_equal(out[key], ref[key], equal_nan=True), (rank, int(s), key)
        t_whole.append(t1 - t0)
        t_split.append(t3 - t2)
        checked += 1
    dist.barrier()
    out = {
        "check": "split controller step (i2lqr_split_*, lap rows exchanged with ncclAllGather) == single-GPU i2lqr_solve_step, bit for bit, outputs and per-candidate trace",
        "bit_identical": True, "world_size": world, "steps_checked": checked,
        "K": int(g["K"]), "L_parity": int(g["L"]), "L_latency": 8, "row_doubles": split8.row, "allgather_bytes_per_rank": int(split8.rows.numel() * 8),
        "p50_ms_single_gpu": round(1e3 * float(np.median(t_whole)), 3), "p50_ms_split": round(1e3 * float(np.median(t_split)), 3),
    }  # fmt: skip
    for e in (whole, part, whole8, part8):
        e.close()
    return out


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    out = run_check(rank, world, local)
    if rank == 0:
        print(json.dumps(out))
        print("OK")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
