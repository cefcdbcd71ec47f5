import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import numpy as np
from judo_b200.engine import Engine
from judo_b200.spline import spline_basis
from judo_b200.tasks.leap_cube import QPOS_HOME, reduced_collision_model
from oracle import plan as op
from oracle.mjc import OracleModel, load_table

tb = load_table("leap_cube")
geoms, pairs = reduced_collision_model(tb)
om = OracleModel(tb, pairs=pairs, geoms=geoms)
rng = np.random.default_rng(5)
N, H, K = 48, 40, 4
eng = Engine("leap_cube", N)
x0 = np.concatenate([QPOS_HOME, np.zeros(22)])
lo = np.array([a["ctrlrange"][0] for a in tb["actuators"]]); hi = np.array([a["ctrlrange"][1] for a in tb["actuators"]])
nominal = np.tile(QPOS_HOME[7:], (K, 1))
knots = np.clip(nominal + 0.2 * 4.0 * np.linspace(0.25, 1, K)[:, None] * rng.normal(size=(N, K, 16)), lo, hi)
times = np.linspace(0, 0.4, K); query = 0.01 * np.arange(H)
ctrl = op.make_spline(times, knots, "cubic")(query)
s_ref, _ = om.rollout(x0, ctrl)
s_gpu, _ = eng.rollout(x0, ctrl)
err = np.abs(s_gpu - s_ref).max(axis=2)
print("per-rollout max err:", np.round(np.log10(err.max(axis=1) + 1e-300), 1))
bad = int(np.argmax(err.max(axis=1)))
print("worst rollout", bad, "err by step", err[bad])
t0 = int(np.argmax(err[bad] > 1e-7))
print("first step with err>1e-7:", t0)
for t in range(max(0, t0 - 2), min(H, t0 + 2)):
    q = s_ref[bad, t - 1, :23] if t > 0 else x0[:23]; v = s_ref[bad, t - 1, 23:] if t > 0 else x0[23:]
    f = om.forward(q, v, ctrl[bad, t])
    print("t", t, "ncon", f["ncon"], "nefc", f["nefc"], "iters", f["solver_iter"], "dist", np.round(f["contact_dist"], 6))
    # one-step comparison from the same (oracle) state
    eng.update(1)
    s1, _ = eng.rollout(np.concatenate([q, v]), ctrl[bad:bad + 1, t:t + 1])
    r1, _ = om.rollout(np.concatenate([q, v]), ctrl[bad:bad + 1, t:t + 1])
    print("   one-step err from oracle state (cold warmstart both):", np.abs(s1 - r1).max())
