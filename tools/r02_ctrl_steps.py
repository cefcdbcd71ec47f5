"""A few Controller.update_action() calls at C2 (for `ncu -k regex:rollout_kernel -s 20 -c 1`: the fused kernel as the Controller launches it,
with elite traces)."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
from judo_b200.controller import make_controller
np.random.seed(42)
c = make_controller("cartpole", "mppi")
c.optimizer_cfg.num_rollouts = 4096
c.controller_cfg.horizon = 2.56
c.reset()
for i in range(30):
    c.time = c.task.dt * i
    c.update_action()
c.engine.close()
