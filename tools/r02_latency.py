"""update_action() latency at the reference's size (C1) and at the bench sizes, with the C call's internal breakdown (B200MPC_TIMING)."""
import os, sys, time, statistics
os.environ["B200MPC_TIMING"] = "1"
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
from judo_b200.controller import make_controller

def run(task, opt, N, horizon, n=200, fast=True, traces=None):
    np.random.seed(42)
    c = make_controller(task, opt)
    c.optimizer_cfg.num_rollouts = N
    c.controller_cfg.horizon = horizon
    c.fast_path = fast
    if traces is not None:
        c.controller_cfg.max_num_traces = traces
    c.reset()
    if hasattr(c.task, "get_sim_metadata"):
        c.system_metadata = c.task.get_sim_metadata()
    for _ in range(20):
        c.update_action()
    lat = []
    for i in range(n):
        c.time = c.task.dt * i
        t = time.perf_counter(); c.update_action(); lat.append(time.perf_counter() - t)
    print(f"{task}+{opt} N={N} H={c.num_timesteps} fast={fast} traces={c.controller_cfg.max_num_traces}: p50 {statistics.median(lat)*1e3:.4f} ms  p10 {np.percentile(lat,10)*1e3:.4f}  p90 {np.percentile(lat,90)*1e3:.4f}", flush=True)
    c.engine.close()

run("cartpole", "ps", 32, 1.28)
run("cartpole", "ps", 32, 1.28, fast=False)
run("cartpole", "mppi", 4096, 2.56)
run("cartpole", "mppi", 4096, 2.56, traces=0)
run("cartpole", "mppi", 4096, 2.56, fast=False)
run("cylinder_push", "cem", 2048, 1.0)
run("leap_cube", "mppi", 1024, 0.4, n=15)
