"""CPU: host-side logic of the product (no GPU compute) and the C-ABI surface."""
import ctypes
import os
import re

import numpy as np
import pytest

from judo_b200 import _lib
from judo_b200.config import OverridableConfig, set_config_overrides
from judo_b200.consts import task_consts
from judo_b200.normalization import make_normalizer
from judo_b200.optimizers import (MPPI, CrossEntropyMethod, CrossEntropyMethodConfig, MPPIConfig, PredictiveSamplingConfig,
                                  get_registered_optimizers, register_optimizer)
from judo_b200.spline import spline_basis

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_spline_basis_matches_reference_golden(golden):
    g = golden("spline")
    for i in range(int(g["ncases"])):
        kind = str(g[f"s{i}_kind"])
        for q, o in (("query", "out"), ("query_shift", "out_shift")):
            B = spline_basis(g[f"s{i}_times"], g[f"s{i}_{q}"], kind)
            np.testing.assert_allclose(np.einsum("hk,nkj->nhj", B, g[f"s{i}_knots"]), g[f"s{i}_{o}"], rtol=0, atol=1e-13)


@pytest.mark.parametrize("kind", ["zero", "linear", "cubic"])
def test_spline_basis_matches_scipy_on_coincident_and_outside_queries(kind):
    from scipy.interpolate import interp1d

    rng = np.random.default_rng(1)
    for K in (4, 5, 9):
        t = 2.0 + np.linspace(0, 1.3, K)
        y = rng.normal(size=(K, 2))
        q = np.concatenate([t, t - 1e-12, t + 1e-12, [t[0] - 1, t[-1] + 1], rng.uniform(t[0], t[-1], 20)])
        ref = interp1d(t, y, kind=kind, axis=0, fill_value=(y[0], y[-1]), bounds_error=False)(q)
        np.testing.assert_allclose(spline_basis(t, q, kind) @ y, ref, rtol=0, atol=1e-10)
    with pytest.raises(ValueError):
        spline_basis(np.arange(3.0), np.arange(3.0), "cubic")
    with pytest.raises(ValueError):
        spline_basis(np.arange(4.0), np.arange(3.0), "quintic")


def test_sampling_is_bit_exact_with_reference(golden, temp_np_seed):
    g = golden("optimizers")
    reg = get_registered_optimizers()
    for ci in range(int(g["ncases"])):
        name, task, nu, N = g[f"c{ci}_meta"][:4]
        cls, cfg_cls = reg[str(name)]
        cfg = cfg_cls()
        if task != "default":
            cfg.set_override(str(task))
        cfg.num_rollouts = int(N)
        opt = cls(cfg, int(nu))
        with temp_np_seed(7 + ci):
            for it in range(2):
                if name == "cem":
                    np.testing.assert_array_equal(opt.sigma, g[f"c{ci}_sigma_in{it}"])
                knots = opt.sample_control_knots(g[f"c{ci}_nominal_in{it}"])
                np.testing.assert_array_equal(knots, g[f"c{ci}_knots{it}"])
                np.testing.assert_array_equal(knots[0], g[f"c{ci}_nominal_in{it}"])  # row 0 is the un-noised nominal
                if name == "cem":
                    opt.sigma = g[f"c{ci}_sigma_out{it}"]


def test_update_without_engine_fails_loudly():
    opt = MPPI(MPPIConfig(), 2)
    with pytest.raises(RuntimeError, match="GPU"):
        opt.update_nominal_knots(np.zeros((16, 4, 2)), np.zeros(16))


def test_cem_pre_optimization_reinterpolates_sigma():
    cfg = CrossEntropyMethodConfig()
    opt = CrossEntropyMethod(cfg, 2)
    opt.sigma = np.arange(8.0).reshape(4, 2)
    cfg.num_nodes = 6
    opt.pre_optimization(np.linspace(0, 1, 4), np.linspace(0.1, 1.1, 6))
    from scipy.interpolate import interp1d

    ref = interp1d(np.linspace(0, 1, 4), np.arange(8.0).reshape(4, 2), axis=0, fill_value="extrapolate", kind="linear")(np.linspace(0.1, 1.1, 6))
    np.testing.assert_allclose(opt.sigma, ref, atol=1e-13)


def test_config_overrides_and_registry():
    cfg = MPPIConfig()
    cfg.set_override("leap_cube")
    assert (cfg.sigma, cfg.temperature, cfg.noise_ramp, cfg.num_rollouts, cfg.use_noise_ramp) == (0.2, 0.0025, 4.0, 32, True)
    cfg.set_override("nonexistent")
    assert (cfg.sigma, cfg.temperature, cfg.noise_ramp, cfg.num_rollouts, cfg.use_noise_ramp) == (0.1, 0.05, 2.5, 16, False)
    ps = PredictiveSamplingConfig()
    ps.set_override("cartpole")
    assert ps.num_rollouts == 32 and ps.sigma == 0.05
    with pytest.warns(UserWarning):
        set_config_overrides("x", MPPIConfig, {"not_a_field": 1})
    with pytest.raises(TypeError):
        set_config_overrides("x", int, {})

    class MyOpt(MPPI):
        pass

    register_optimizer("mine", MyOpt, MPPIConfig)
    assert get_registered_optimizers()["mine"] == (MyOpt, MPPIConfig)
    assert issubclass(MPPIConfig, OverridableConfig)


def test_normalizers_closed_form():
    mm = make_normalizer("min_max", 2, min=np.array([-2.0, 0.0]), max=np.array([2.0, 4.0]))
    x = np.array([[0.0, 1.0], [2.0, 4.0]])
    np.testing.assert_allclose(mm.normalize(x), [[0, -0.5], [1, 1]])
    np.testing.assert_allclose(mm.denormalize(mm.normalize(x)), x)
    with pytest.warns(UserWarning):
        inf = make_normalizer("min_max", 2, min=np.array([-np.inf, 0.0]), max=np.array([np.inf, 1.0]))
    np.testing.assert_allclose(inf.normalize(np.array([3.0, 0.5])), [3.0, 0.0])
    rn = make_normalizer("running", 3, init_std=1.0)
    data = np.random.default_rng(0).normal(size=(5, 7, 3)) * 3 + 1
    for chunk in data:
        rn.update(chunk)
    np.testing.assert_allclose(rn.mean, data.reshape(-1, 3).mean(0), rtol=1e-12)
    np.testing.assert_allclose(rn.std, data.reshape(-1, 3).std(0), rtol=1e-12)
    np.testing.assert_allclose(rn.denormalize(rn.normalize(data)), data, rtol=1e-4, atol=1e-5)  # eps only on the way in
    assert make_normalizer("none", 2).normalize(x) is x
    with pytest.raises(ValueError):
        make_normalizer("bogus", 2)


def test_task_constant_tables_have_the_struct_sizes():
    # sizes of the all-double structs in csrc/small_tasks.cuh, csrc/fr3.cuh
    assert task_consts("cartpole").size == 38
    assert task_consts("cylinder_push").size == 39
    assert task_consts("fr3_pick").size == 592


def test_fr3_pick_task_surface_and_phase_machine(golden):
    """Host side of the fr3_pick task (no GPU): indices as the reference's __init__ derives them (fr3_pick.py:111-144), the phase
    machine against the reference's own pre_rollout outputs, the inherited control range of the gripper servo."""
    from judo_b200.tasks.fr3_pick import QPOS_HOME, FR3Pick, Phase

    task = FR3Pick()
    assert (task.model.nq, task.model.nv, task.model.nu, task.model.nsensordata) == (16, 15, 8, 14)
    assert task.obj_pos_adr == 0 and task.arm_pos_slice == slice(7, 16) and task.dt == 0.004
    assert (task.left_finger_table_adr, task.right_finger_table_adr, task.obj_table_adr, task.ee_z_adr, task.grasp_site_adr) == (2, 3, 4, 5, 11)
    np.testing.assert_array_equal(task.data.qpos, QPOS_HOME)
    np.testing.assert_allclose(task.actuator_ctrlrange[7], [-0.02, 0.06])  # inheritrange="2.0" on a 0..0.04 joint
    np.testing.assert_allclose(task.actuator_ctrlrange[3], [-3.0421, -0.1518])
    g = golden("rewards_fr3")
    for x, ph in zip(g["fr3_phase_states"], g["fr3_phases"]):
        task.pre_rollout(x)
        assert task.phase == Phase(int(ph))
        assert task.cost_params()[0] == float(ph) and task.cost_params().size == 23


def test_c_abi_library_exports_every_declared_symbol():
    """include/b200mpc.h, judo_b200/_lib.py and the built library agree (loading needs no GPU)."""
    header = open(os.path.join(ROOT, "include", "b200mpc.h")).read()
    declared = set(re.findall(r"\b(b200mpc_[a-z0-9_]+)\s*\(", header))
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    lib = _lib.load()
    for name in declared:
        assert hasattr(lib, name)
    # no device here: create must fail cleanly with a message, never crash or fall back
    if not os.path.exists("/dev/nvidia0"):
        h = ctypes.c_void_p()
        c = task_consts("cartpole")
        rc = lib.b200mpc_create(ctypes.byref(h), 0, c.ctypes.data, c.size, 0, 8)
        assert rc != 0 and b"CUDA" in lib.b200mpc_last_error(None)


def test_leap_pair_list_is_mujocos_static_filter_in_mj_collision_order():
    """The leap_cube collision pairs the oracle and the kernel share: every geom pair MuJoCo's static filters leave (same body, parent-child
    unless one side is welded to the world, the 18 <exclude>s of leap_components/params_and_default.xml:76-101, contype/conaffinity), the
    cube's 71 pairs first, then the hand-hand pairs body pair by body pair."""
    from judo_b200.tasks.leap_cube import LEAP_MAXHH, leap_consts, reduced_collision_model
    from oracle.mjc import load_table

    tb = load_table("leap_cube")
    geoms, pairs = reduced_collision_model(tb)
    names = [b["name"] for b in tb["bodies"]]
    body = lambda g: names[geoms[g]["body"]]  # noqa: E731
    cube = next(i for i, g in enumerate(geoms) if g["name"] == "cube")
    assert len(pairs) == 1692 and all(cube in p for p in pairs[:71]) and not any(cube in p for p in pairs[71:])
    hh = pairs[71:]
    assert len(hh) == 1621 <= LEAP_MAXHH and len({tuple(p) for p in hh}) == len(hh)
    keys = [(geoms[a]["body"], geoms[b]["body"], a, b) for a, b in hh]
    assert keys == sorted(keys)                                    # body-pair major, geoms of the first body outermost
    bp = {(body(a), body(b)) for a, b in hh}
    assert len(bp) == 106
    excluded = [("palm", "if_bs"), ("palm", "if_px"), ("palm", "if_md"), ("palm", "th_px"), ("if_bs", "mf_bs"), ("th_mp", "rf_bs")]
    for a, b in excluded:
        assert (a, b) not in bp and (b, a) not in bp
    for a, b in [("if_bs", "if_px"), ("mf_md", "mf_ds"), ("palm", "palm")]:   # parent-child (neither welded to the world) / same body
        assert (a, b) not in bp and (b, a) not in bp
    for a, b in [("palm", "if_ds"), ("if_bs", "if_md"), ("if_md", "mf_md"), ("rf_ds", "th_ds")]:  # fingertip-palm, grandparent, neighbours
        assert (a, b) in bp
    assert {g["type"] for g in geoms} == {"box", "sphere"}          # the four fingertip meshes are the only substitution
    c = leap_consts(tb)
    codes = c[-LEAP_MAXHH // 4:].view(np.uint16)
    hidx = {gi: k for k, gi in enumerate(i for i in range(len(geoms)) if i != cube)}
    assert [int(x) for x in codes[:len(hh)]] == [hidx[a] * 256 + hidx[b] for a, b in hh] and not codes[len(hh):].any()


@pytest.mark.parametrize("kind,size", [("capsule", [0.045, 0.5, 0.0]), ("capsule", [0.1, 0.05, 0.0]), ("cylinder", [0.25, 0.05, 0.0]),
                                       ("sphere", [0.3, 0.0, 0.0]), ("box", [0.03, 0.05, 0.07])])
def test_inertia_from_geom_matches_numerical_quadrature(kind, size):
    """An external referee for the MJCF compiler's solids (the cartpole pole is a capsule whose inertia comes from its geometry,
    cartpole.xml:28-30): mass and principal inertia of every primitive against a midpoint-rule integral over the solid itself."""
    from judo_b200.mjcf import _geom_mass_inertia

    n = 120
    if kind == "capsule":
        r, hh = size[0], size[1]
        L = [r, r, hh + r]
        inside = lambda x, y, z: ((np.hypot(x, y) <= r) & (np.abs(z) <= hh)) | (x * x + y * y + (np.abs(z) - hh).clip(0) ** 2 <= r * r)  # noqa: E731
    elif kind == "cylinder":
        r, hh = size[0], size[1]
        L = [r, r, hh]
        inside = lambda x, y, z: (np.hypot(x, y) <= r) & (np.abs(z) <= hh)  # noqa: E731
    elif kind == "sphere":
        r = size[0]
        L = [r, r, r]
        inside = lambda x, y, z: x * x + y * y + z * z <= r * r  # noqa: E731
    else:
        L = list(size)
        inside = lambda x, y, z: np.ones_like(x, dtype=bool)  # noqa: E731
    ax = [(np.arange(n) + 0.5) / n * 2 * l - l for l in L]
    X, Y, Z = np.meshgrid(*ax, indexing="ij")
    m = inside(X, Y, Z)
    dv = np.prod([2 * l / n for l in L])
    vol = m.sum() * dv
    per_mass = np.array([((Y * Y + Z * Z) * m).sum(), ((X * X + Z * Z) * m).sum(), ((X * X + Y * Y) * m).sum()]) * dv / vol
    mass, inertia = _geom_mass_inertia(kind, np.array(size, dtype=float), 1000.0, None)
    assert abs(mass / 1000.0 - vol) / vol < 2e-3
    np.testing.assert_allclose(inertia / mass, per_mass, rtol=2e-3)
    m2, i2 = _geom_mass_inertia(kind, np.array(size, dtype=float), 1000.0, 0.1)   # explicit mass (the pole: mass="0.1") scales the same solid
    assert m2 == 0.1
    np.testing.assert_allclose(i2 / m2, inertia / mass, rtol=1e-12)
