"""TEST INFRASTRUCTURE: an Engine whose kernels run on the CPU SIMT emulator (tests/warpsim) instead of the GPU, so that the
no-GPU tier can drive judo_b200.controller.Controller end to end (time shift, sampling, clip, fused plan step, traces) against the
reference Controller's golden plan steps.  It executes the SAME device code as libb200mpc.so (compiled with g++ -DB2_HOST_SIM);
what it does NOT execute is the host side of the C ABI (buffer management, launches) and, for the warp-per-rollout tasks, the
stand-alone reduction kernels of the optimizer update (restated here with oracle.plan; the GPU tier covers both).
Never imported by the product."""
from __future__ import annotations

import ctypes

import numpy as np

from judo_b200.consts import task_consts
from judo_b200.engine import Engine
from oracle import plan as op
from tests import warpsim

P = lambda a: None if a is None else a.ctypes.data_as(ctypes.c_void_p)  # noqa: E731
_DIMS = {"cartpole": (2, 2, 1, 6, 6), "cylinder_push": (4, 4, 2, 6, 6), "leap_cube": (23, 22, 16, 31, 9), "fr3_pick": (16, 15, 8, 14, 23)}
_OPT = {"mppi": 0, "cem": 1, "ps": 2}


class SimEngine(Engine):
    supports_controller_step = False  # the emulator covers the device code; the C-side host glue has its own CPU tests
    def __init__(self, task: str, num_rollouts: int, device: int = 0, consts: np.ndarray | None = None) -> None:  # noqa: ARG002
        self._sim = warpsim.lib()
        self.task = task
        self._consts = np.ascontiguousarray(task_consts(task) if consts is None else consts, dtype=np.float64)
        self.nq, self.nv, self.nu, self.nsensordata, self.n_cost_params = _DIMS[task]
        self._N = int(num_rollouts)
        self._capture, self._trace = False, None
        self._h = None
        self.device = device
        self.launches = 0

    def close(self) -> None: ...

    @property
    def num_rollouts(self) -> int:
        return self._N

    @property
    def launch_count(self) -> int:
        return self.launches

    @property
    def contact_overflows(self) -> int:
        return 0

    def update(self, num_rollouts: int) -> None:
        self._N = int(num_rollouts)

    @property
    def trace_width(self) -> int:
        return {"leap_cube": 15, "fr3_pick": 6}.get(self.task, 0)

    def set_trace_capture(self, enable: bool) -> None:
        self._capture = bool(enable)

    def elite_traces(self, idx: np.ndarray, H: int) -> np.ndarray:
        assert self._trace is not None and self._trace.shape[1] == H
        return self._trace[np.asarray(idx)].copy()

    def rollout(self, x0: np.ndarray, controls: np.ndarray, want_sensors: bool = True):  # noqa: ANN201
        x0, controls = np.ascontiguousarray(x0, dtype=np.float64), np.ascontiguousarray(controls, dtype=np.float64)
        N, H, _ = controls.shape
        if N != self._N:
            raise RuntimeError("controls batch size does not match num_rollouts (call update first)")
        s, e = np.zeros((N, H, self.nq + self.nv)), np.zeros((N, H, self.nsensordata))
        b = int(x0.ndim == 2)
        if self.task in ("cartpole", "cylinder_push"):
            self._sim.sim_rollout(0 if self.task == "cartpole" else 1, P(self._consts), P(x0), b, P(controls), N, H, P(s), P(e), 32)
        elif self.task == "leap_cube":
            self._sim.sim_leap_rollout(P(self._consts), P(x0), b, P(controls), N, H, P(s), P(e), 2, 3, 0)
        else:
            self._sim.sim_fr3_rollout(P(self._consts), P(x0), b, P(controls), N, H, P(s), P(e), 2, 3, 0)
        self.launches += 1
        return s, (e if want_sensors else None)

    def plan_step(self, x0, knots, basis, cost_params, optimizer, opt_params, want_rewards=True, n_elite=0):  # noqa: ANN001, ANN201
        x0, knots, basis = (np.ascontiguousarray(a, dtype=np.float64) for a in (x0, knots, basis))
        params = np.ascontiguousarray(cost_params, dtype=np.float64)
        N, K, nu = knots.shape
        H = basis.shape[0]
        opl = [float(v) for v in np.atleast_1d(opt_params)] if opt_params is not None and np.size(opt_params) else []
        rew, nom, sig, el = np.zeros(N), np.zeros((K, nu)), np.zeros((K, nu)), np.full(8, -1.0)
        self.launches += 1
        if self.task in ("cartpole", "cylinder_push"):
            opp = np.ascontiguousarray(opl + [0.0])
            self._sim.sim_plan_step(0 if self.task == "cartpole" else 1, P(self._consts), P(x0), P(knots), N, K, P(basis), H, P(params),
                                    _OPT[optimizer], P(opp), int(n_elite), 32, None, P(rew), P(nom), P(sig), P(el), 0)
            return dict(nominal=nom, sigma=sig if optimizer == "cem" else None, rewards=rew, elite=el[:n_elite].astype(np.int32))
        nt = self.trace_width
        self._trace = np.zeros((N, H, nt)) if self._capture else None
        fn = self._sim.sim_leap_plan_costs if self.task == "leap_cube" else self._sim.sim_fr3_plan_costs
        fn(P(self._consts), P(x0), P(knots), N, K, P(basis), H, P(params), None, P(rew), 2, 3, 0, P(self._trace))
        # optimizer update + elite list of the warp-per-rollout tasks: stand-alone reduction kernels in the product, oracle.plan here
        if optimizer == "mppi":
            nom = op.mppi_update(knots, rew, opl[0])
        elif optimizer == "cem":
            nom, sig = op.cem_update(knots, rew, int(opl[0]), opl[1], opl[2])
        else:
            nom = op.ps_update(knots, rew)
        elite = np.argsort(rew, kind="stable")[::-1][:n_elite]
        return dict(nominal=nom, sigma=sig if optimizer == "cem" else None, rewards=rew, elite=elite.astype(np.int32))
