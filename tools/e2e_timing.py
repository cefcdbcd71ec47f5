import os, sys, time
os.environ["B200MPC_TIMING"] = "1"
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import numpy as np
import bench
from judo_b200.engine import Engine
w = dict(bench.WORKLOADS["cartpole_mppi"])
task, opt, x0, knots, basis, params, _ = bench.problem(w, 4096)
for zc in ("0", "1", "2", "3"):
    os.environ["B200MPC_ZEROCOPY"] = zc
    eng = Engine("cartpole", 4096)
    op = opt.fused_params()
    for _ in range(50): eng.plan_step(x0, knots, basis, params, "mppi", op, True, 5)
    t = []
    for _ in range(300):
        t0 = time.perf_counter(); eng.plan_step(x0, knots, basis, params, "mppi", op, True, 5); t.append(time.perf_counter() - t0)
    print("zerocopy", zc, "python-level p50 us", np.median(t) * 1e6, "min", np.min(t) * 1e6)
    eng.close()
