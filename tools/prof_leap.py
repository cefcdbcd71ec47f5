import os, sys
os.environ["B200MPC_LEAP_PROF"] = "1"
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import numpy as np
import bench
from judo_b200.engine import Engine
w = dict(bench.WORKLOADS["leap_cube_mppi"])
task, opt, x0, knots, basis, params, _ = bench.problem(w, 1024)
eng = Engine("leap_cube", 1024)
r, _ = eng.plan_costs(x0, knots, basis, params)
print("rewards", r[:3])
eng.close()
