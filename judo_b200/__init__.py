"""judo_b200 — B200-native rollout engine behind judo's sampling-MPC plugin surface.

Public surface (mirrors the reference's names):
    judo_b200.optimizers.{MPPI, CrossEntropyMethod, PredictiveSampling, get_registered_optimizers, ...}
    judo_b200.tasks.{Task, Cartpole, CylinderPush, LeapCube, get_registered_tasks, ...}
    judo_b200.rollout_backend.{RolloutBackend, B200RolloutBackend}
    judo_b200.controller.{Controller, ControllerConfig, make_controller, make_spline}
    judo_b200.engine.Engine          numpy-facing wrapper over the C ABI (include/b200mpc.h)
All compute runs in libb200mpc.so (hand-written sm_100a CUDA); importing the engine without it raises.
"""

__version__ = "0.1.0"
