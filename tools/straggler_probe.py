import os, sys, time, json
import numpy as np
sys.path.insert(0, "/root/repo")
from ilqr_iterative_tasks_b200 import _abi, engine, episodes, synth
sc = synth.episode_scenarios(4096)
p = _abi.default_params(num_ss_points=8, num_ss_iter=2)
lap0, _ = synth.initial_lap()
obs = _abi.obstacle_array(sc["obs"], sc["mode"])
eb = episodes.EpisodeBatch(p, sc["x0"], obstacles=obs, initial_lap=lap0, device=0)
lens = eb.run_lap(); st = eb.status()
bad = np.flatnonzero(st != 0)
print("lap1 stragglers", len(bad), "status counts", np.bincount(st), file=sys.stderr)
lens = eb.run_lap(); st2 = eb.status()
bad2 = np.flatnonzero(st2 != 0)
print("lap2 stragglers", len(bad2), file=sys.stderr)
eb.close()
for name, idx in (("lap2-stragglers", bad2), ("30 of them", bad2[:30])):
    eb = episodes.EpisodeBatch(p, sc["x0"][idx], obstacles=obs[idx], initial_lap=lap0, device=0)
    for lap in range(2):
        t0 = time.perf_counter(); lens = eb.run_lap(); w = time.perf_counter() - t0
        s = eb.stats()
        sp = eb.engine.spec_stats()
        print({k: sp[k] for k in ("hits", "waited", "self_solved", "misses", "spec_finished", "spec_abandoned", "launches")}, file=sys.stderr)
        print(name, "E", len(idx), "lap", lap, "wall %.3f" % w, "steps max", lens.max(), "mean %.1f" % lens.mean(), "iters/solve %.1f" % (s["ilqr_iterations"] / max(1, s["solves"])), "ms/step %.2f" % (1e3 * w / lens.max()), file=sys.stderr)
    eb.close()
