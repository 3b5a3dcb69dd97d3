import sys, time, numpy as np
sys.path.insert(0,'/root/repo')
from ilqr_iterative_tasks_b200 import _abi, engine, episodes, synth
p = _abi.default_params(num_ss_points=64, num_ss_iter=8)
lap0,_ = synth.initial_lap()
obs = _abi.obstacle_array(np.array([[31.0,-3.0,8.0,6.0,0.0]]),[0])
eb = episodes.EpisodeBatch(p, np.zeros((1,4)), obstacles=obs, initial_lap=lap0, device=0)
laps=[lap0]
for _ in range(8):
    n,lx,_ = eb.run_lap(want_traj=True); laps.append(lx[0,:int(n[0])+1].copy())
eb.close()
eng = engine.Engine(p, device=0, n_episodes=1, lap_capacity=128)
for lp in laps[-8:]: eng.add_lap(0, lp)
eng.set_obstacle(0, _abi.Obstacle(31.0,-3.0,8.0,6.0,0.0,0,1))
states = laps[-1][:-6]
rows=[]
for rep in range(3):
    for x in states:
        t1=time.perf_counter(); eng.solve_step(x.reshape(1,4)); w=(time.perf_counter()-t1)*1e3
        st=eng.last_stats(); tr=eng.step_trace()
        it=tr['cand_iters'][0]  # (3,L,K)
        chain=np.where(it>0,it,0).sum(axis=2)  # (3,L)
        rows.append((w, st['kernel_ms'], st['ilqr_iterations'], chain.max(axis=1).sum(), chain.sum()))
r=np.array(rows[len(states):])
print('wall p50 %.3f  device p50 %.3f  iters/step %.0f  critical-chain iters/step %.0f (sum over outer of max chain)  total/8 %.0f' % (np.median(r[:,0]), np.median(r[:,1]), r[:,2].mean(), r[:,3].mean(), r[:,4].mean()/8))
print('us per critical iteration: %.2f' % (1e3*np.median(r[:,1])/r[:,3].mean()))
