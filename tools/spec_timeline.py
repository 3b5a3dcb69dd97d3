"""Developer aid: where a one-launch controller step spends its time (run with I2LQR_SPEC_TIMELINE=1 on a GPU box).
Config-2 shape (K=64, L=8, stock scenario), prints per-outer-iteration phase durations of the slowest chain, medians over steps."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("I2LQR_SPEC_TIMELINE", "1")
from ilqr_iterative_tasks_b200 import _abi, engine, episodes, synth  # noqa: E402

p = _abi.default_params(num_ss_points=64, num_ss_iter=8)
lap0, _ = synth.initial_lap()
obs = _abi.obstacle_array(np.array([[31.0, -3.0, 8.0, 6.0, 0.0]]), [0])
eb = episodes.EpisodeBatch(p, np.zeros((1, 4)), obstacles=obs, initial_lap=lap0)
laps = [lap0]
for _ in range(8):
    n, lx, _ = eb.run_lap(want_traj=True)
    laps.append(lx[0, : int(n[0]) + 1].copy())
eb.close()
eng = engine.Engine(p, n_episodes=1, lap_capacity=128)
for lp in laps[-8:]:
    eng.add_lap(0, lp)
eng.set_obstacle(0, _abi.Obstacle(31.0, -3.0, 8.0, 6.0, 0.0, 0, 1))
rows, items = [], []
for rep in range(2):
    for x in laps[-1][:-6]:
        eng.solve_step(x.reshape(1, 4))
        if rep == 0:
            continue
        tl = eng.spec_timeline().astype(np.int64)[:, :8]  # (3, 8 chains, 8)
        dev = eng.last_stats()["kernel_ms"] * 1e3
        t0 = tl[0, :, 0].min()
        per_it = []
        for it in range(3):
            c = tl[it]
            start, knn, first, walk, rel = [c[:, i] for i in range(5)]
            slow = int(np.argmax(walk))
            per_it.append([
                (knn - start).max() / 1e3, (first - knn).max() / 1e3, (walk - first)[slow] / 1e3, (rel - walk).min() / 1e3,
                (rel.max() - start.min()) / 1e3, c[slow, 5] / 1e3, c[slow, 6] >> 32, c[slow, 6] & 0xFFFFFFFF, (c[:, 6] >> 32).sum(),
            ])
        rows.append((dev, (tl[2, :, 4].max() - t0) / 1e3, per_it))
        dbg = eng.spec_item_debug.astype(np.int64)
        dbg = dbg[: (dbg.size // 4) * 4].reshape(-1, 4)
        ok = (dbg[:, 1] > dbg[:, 0]) & (dbg[:, 0] > 0) & (dbg[:, 2] > 0)
        items.append(np.column_stack([(dbg[ok, 0] - t0) / 1e3, (dbg[ok, 1] - dbg[ok, 0]) / 1e3, dbg[ok, 2], dbg[ok, 3]]))
dev = np.median([r[0] for r in rows])
span = np.median([r[1] for r in rows])
per = np.median(np.array([r[2] for r in rows]), axis=0)
print("device step p50 %.0f us (events), first resolver start -> last release %.0f us" % (dev, span))
print("per outer iteration (medians over %d steps), us:" % len(rows))
print("  it | knn | rank-0 solve (max over chains) | walk of the slowest chain | winner | total | slowest chain: waiting, in-place solves, their iterations | in-place solves all chains")
for it in range(3):
    print("  %d  | %4.0f | %6.0f | %6.0f | %4.0f | %6.0f | %6.0f %3.0f %4.0f | %3.0f" % ((it,) + tuple(per[it])))
it_all = np.concatenate(items)
print("speculative entries finished: %d; start after step start p50 %.0f us p90 %.0f; us per round p50 %.2f p90 %.2f; iterations per round %.2f"
      % (len(it_all), np.median(it_all[:, 0]), np.percentile(it_all[:, 0], 90), np.median(it_all[:, 1] / it_all[:, 2]),
         np.percentile(it_all[:, 1] / it_all[:, 2], 90), it_all[:, 3].sum() / it_all[:, 2].sum()))
wd = eng.spec_worker_debug.astype(np.float64)
wd = wd[wd[:, 0] > 0]
print("worker warps of the last step: %d; cycles per round: picking entries %.0f, round itself %.0f, publishing %.0f (rounds per warp p50 %.0f)"
      % (len(wd), np.median(wd[:, 1] / wd[:, 0]), np.median(wd[:, 2] / wd[:, 0]), np.median(wd[:, 3] / wd[:, 0]), np.median(wd[:, 0])))
tp = eng.spec_trip_profile.astype(np.float64)
tp = tp[tp[:, 4] > 0]
if len(tp):
    print("inside a round (profile build), cycles: terms %.0f, riccati %.0f, forward %.0f, accept %.0f" % tuple(np.median(tp[:, i] / tp[:, 4]) for i in range(4)))
long = it_all[it_all[:, 3] >= 60]
if len(long):
    print("entries with >= 60 iterations: %d, duration p50 %.0f us, rounds p50 %.0f, us per round %.2f" % (len(long), np.median(long[:, 1]), np.median(long[:, 2]), np.median(long[:, 1] / long[:, 2])))
eng.close()
