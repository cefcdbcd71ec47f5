"""Newton-iteration imbalance study for the leap kernel (oracle side; numbers in profiles/r02_leap_barrier_study.md)."""
import sys, ctypes, numpy as np
sys.path.insert(0, '' + __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
import bench
from oracle.mjc import lib
w = dict(bench.WORKLOADS["leap_cube_mppi"])
task, opt, x0, knots, basis, params, _ = bench.problem(w, 1024)
om = bench._oracle_model("leap_cube")
controls = np.einsum("hk,nkj->nhj", basis, knots)
N, H = 1024, 40
stats = np.zeros((N, H, 4))
lib().mjc_set_stats_buffer.argtypes = [ctypes.c_void_p]
lib().mjc_set_stats_buffer(stats.ctypes.data)
om.rollout(x0, controls)
lib().mjc_set_stats_buffer(None)
it = stats[:, :, 0]
np.save("/tmp/leap_iters.npy", stats)
print("mean iters", it.mean(), "max", it.max(), "mean ncon", stats[:,:,3].mean(), "max ncon", stats[:,:,3].max())
# cost model: per step per block = A + B*max_iters(block); A = non-solver part ~ equivalent of a_it iterations
def total(groups_per_step, a_it=2.0):
    # groups_per_step[t] = array (nblocks, 7) of rollout ids
    tot = 0.0
    for t in range(H):
        g = groups_per_step[t]
        tot += (a_it + it[g, t].max(axis=1)).max()   # kernel = slowest block? no: blocks are independent across steps unless grid-synced
    return tot
ids = np.arange(N)
pad = (-N) % 7
idp = np.concatenate([ids, np.full(pad, -1)]).reshape(-1, 7)
# static: each block independent, time = max over blocks of sum_t (A + max_w iters)
def static_time(a_it):
    itp = np.concatenate([it, np.zeros((pad, H))])[idp.clip(0)]  # (nb,7,H)
    per_block = (a_it + itp.max(axis=1)).sum(axis=1)
    return per_block.max(), per_block.mean()
def ideal_time(a_it):  # no idle: each warp own pace
    per = (a_it + it).sum(axis=1)
    return per.max(), per.mean()
def sorted_time(a_it, key="prev"):
    tot = 0
    for t in range(H):
        k = it[:, t-1] if t > 0 else np.zeros(N)
        if key == "oracle": k = it[:, t]
        order = np.argsort(k, kind="stable")
        g = np.concatenate([order, np.full(pad, order[-1])]).reshape(-1, 7)
        tot += (a_it + it[g, t].max(axis=1)).max()  # grid sync per step: slowest block
    return tot
for a_it in (1.5, 2.5):
    print("A =", a_it, "iteration-equivalents")
    print("  static  (max block, mean block):", static_time(a_it))
    print("  ideal free-running (max warp, mean warp):", ideal_time(a_it))
    print("  sorted by previous step's count, grid sync per step:", sorted_time(a_it))
    print("  sorted by this step's count (oracle knowledge):", sorted_time(a_it, "oracle"))
# autocorrelation
c = np.corrcoef(it[:, 1:].ravel(), it[:, :-1].ravel())[0, 1]
print("corr(it_t, it_t-1) =", c)
print("per-step max over all rollouts, mean:", it.max(axis=0).mean(), " per-step mean:", it.mean(axis=0).mean())
stats = np.load("/tmp/leap_iters.npy"); it = stats[:, :, 0]; N, H = it.shape
pad = (-N) % 7
itp = np.concatenate([it, np.zeros((pad, H))]).reshape(-1, 7, H)
for a_it in (1.5, 2.5):
    for s in (1, 2, 4, 5, 8, 10, 20, 40):
        per_block = 0
        for t0 in range(0, H, s):
            per_block = per_block + (a_it * min(s, H - t0) + itp[:, :, t0:t0+s].sum(axis=2)).max(axis=1)
        print(f"A={a_it} barrier every {s:2d} steps: max block {per_block.max():.0f} mean block {per_block.mean():.0f}")
# different group sizes at s=1
for wpb in (1, 2, 4, 7):
    pad = (-N) % wpb
    x = np.concatenate([it, np.zeros((pad, H))]).reshape(-1, wpb, H)
    pb = (2.0 + x.max(axis=1)).sum(axis=1)
    print("wpb", wpb, "max", pb.max(), "mean", pb.mean())
