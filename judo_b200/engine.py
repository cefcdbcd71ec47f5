"""Engine: a thin numpy-facing wrapper over one b200mpc handle (one GPU).  All compute happens in libb200mpc.so."""

from __future__ import annotations

import ctypes

import numpy as np

from judo_b200 import _lib
from judo_b200.consts import TASK_IDS, task_consts

OPT_IDS = {"mppi": 0, "cem": 1, "ps": 2}
SPLINE_IDS = {"zero": 0, "linear": 1, "cubic": 2}


class LegacyStream:
    """Direct view of the state of NumPy's GLOBAL legacy generator (the stream behind np.random.randn, which the reference samples its
    candidates from: judo/optimizers/mppi.py:58).  b200mpc_controller_step advances that state itself, bit for bit as numpy would;
    what cannot be reached from outside numpy is the generator's cached second gaussian, so ``head`` draws the first one or two normals
    of a block THROUGH numpy, after which the cache is known to be empty (see include/b200mpc.h)."""

    def __init__(self) -> None:
        self._rs = np.random.mtrand._rand
        self._bg = self._rs._bit_generator
        self.ok = type(self._bg).__name__ == "MT19937"
        if not self.ok:
            return
        addr = self._bg.ctypes.state_address  # numpy's mt19937_state { uint32 key[624]; int pos; }
        self.key_addr, self.pos_addr = addr, addr + 624 * 4
        self._pos = ctypes.c_int.from_address(self.pos_addr)
        self._key0 = ctypes.c_uint32.from_address(addr)
        st = self._rs.get_state(legacy=True)
        key = np.ctypeslib.as_array((ctypes.c_uint32 * 624).from_address(addr))
        self.ok = st[0] == "MT19937" and np.array_equal(key, st[1]) and self._pos.value == st[2]
        self._draw = self._rs.standard_normal

    def current(self) -> bool:
        """False when somebody replaced numpy's global RandomState or its bit generator since this view was made."""
        return self.ok and np.random.mtrand._rand is self._rs and self._rs._bit_generator is self._bg

    def head(self, n: int) -> list[float]:
        """The first min(n, 1 or 2) normals of a block of n, drawn through numpy so that its gaussian cache ends up empty (n >= 2)."""
        if n <= 0:
            return []
        p0, k0 = self._pos.value, self._key0.value
        z0 = self._draw()
        if n == 1:
            return [z0]
        if self._pos.value != p0 or self._key0.value != k0:  # the state advanced: a fresh pair was made, its second value is cached
            return [z0, self._draw()]
        return [z0]                                          # z0 WAS the cached value

    def tail(self) -> float:
        return self._draw()

    def randn(self, n: int) -> np.ndarray:
        """np.random.randn(n), same values and same final generator state, ~3x faster: all but at most three of the normals are
        produced by the library's batched restatement of legacy_gauss (b200mpc_legacy_normals) straight from numpy's state."""
        out = np.empty(n)
        head = self.head(n)
        out[:len(head)] = head
        rem = n - len(head)
        if rem >= 2:
            if _lib.load().b200mpc_legacy_normals(self.key_addr, self.pos_addr, out.ctypes.data + 8 * len(head), rem & ~1):
                raise RuntimeError("b200mpc_legacy_normals failed")
        if rem & 1:
            out[n - 1] = self.tail()
        return out


_stream: LegacyStream | None = None


def legacy_stream() -> LegacyStream:
    """The process-wide view of numpy's global legacy generator (rebuilt if numpy's RandomState object was replaced)."""
    global _stream
    if _stream is None or (_stream.ok and not _stream.current()):
        _stream = LegacyStream()
    return _stream
_dp = ctypes.POINTER(ctypes.c_double)


def _p(a: np.ndarray | None):  # noqa: ANN202
    return None if a is None else a.ctypes.data


def _c(a: np.ndarray) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float64)


class Engine:
    """Owns a b200mpc_handle.  Raises RuntimeError with the library's message on any failure."""

    supports_controller_step = True  # b200mpc_controller_step (the Controller's one-call fast path)
    # While the GPU runs a plan step, draw the NEXT step's normals from a copy of numpy's generator state; the block is used only if the
    # generator is still in exactly that state when the next step starts (include/b200mpc.h) — same stream, sampling time hidden.
    speculative_sampling = True
    SPECULATE_MIN = 2048      # normals per block below which sampling is too cheap to be worth hiding

    def __init__(self, task: str, num_rollouts: int, device: int = 0, consts: np.ndarray | None = None) -> None:
        self._lib = _lib.load()
        self.task = task
        c = _c(task_consts(task) if consts is None else consts)
        h = ctypes.c_void_p()
        rc = self._lib.b200mpc_create(ctypes.byref(h), TASK_IDS[task], _p(c), c.size, device, int(num_rollouts))
        if rc:
            raise RuntimeError("b200mpc_create: " + self._lib.b200mpc_last_error(None).decode())
        self._h = h
        d = _lib.Dims()
        self._check(self._lib.b200mpc_get_dims(self._h, ctypes.byref(d)))
        self.nq, self.nv, self.nu, self.nsensordata, self.n_cost_params = d.nq, d.nv, d.nu, d.nsensordata, d.n_cost_params
        self.device = device

    def close(self) -> None:
        if getattr(self, "_h", None):
            self._lib.b200mpc_destroy(self._h)
            self._h = None

    def __del__(self) -> None:
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass

    def _check(self, rc: int) -> None:
        if rc:
            raise RuntimeError(self._lib.b200mpc_last_error(self._h).decode())

    @property
    def handle(self) -> ctypes.c_void_p:
        return self._h

    @property
    def num_rollouts(self) -> int:
        return self._lib.b200mpc_num_rollouts(self._h)

    @property
    def launch_count(self) -> int:
        return int(self._lib.b200mpc_launch_count(self._h))

    # ---- trace capture (warp-per-rollout tasks)
    @property
    def trace_width(self) -> int:
        """Doubles per step the fused kernel can keep per rollout (0: the task has no trace capture)."""
        return int(self._lib.b200mpc_trace_width(self._h))

    def set_trace_capture(self, enable: bool) -> None:
        self._check(self._lib.b200mpc_set_trace_capture(self._h, int(enable)))

    def elite_traces(self, idx: np.ndarray, H: int) -> np.ndarray:
        """Trace sensors (len(idx), H, trace_width) of the given rollouts of the LAST fused plan step."""
        idx = np.ascontiguousarray(idx, dtype=np.int32)
        out = np.empty((len(idx), H, self.trace_width))
        self._check(self._lib.b200mpc_elite_traces(self._h, idx.ctypes.data, len(idx), int(H), out.ctypes.data))
        return out

    @property
    def contact_overflows(self) -> int:
        """Rollout steps (process-wide) whose contact count exceeded the kernel's per-step buffer (truncation is never silent)."""
        n = int(self._lib.b200mpc_contact_overflows(self._h))
        if n < 0:
            raise RuntimeError(self._lib.b200mpc_last_error(self._h).decode())
        return n

    def update(self, num_rollouts: int) -> None:
        self._check(self._lib.b200mpc_update(self._h, int(num_rollouts)))

    # ---- contract A
    def rollout(self, x0: np.ndarray, controls: np.ndarray, want_sensors: bool = True) -> tuple[np.ndarray, np.ndarray | None]:
        x0, controls = _c(x0), _c(controls)
        N, H, _ = controls.shape
        states = np.empty((N, H, self.nq + self.nv))
        sensors = np.empty((N, H, self.nsensordata)) if want_sensors else None
        self._check(self._lib.b200mpc_rollout(self._h, _p(x0), int(x0.ndim == 2), _p(controls), N, H, _p(states), _p(sensors)))
        return states, sensors

    # ---- contract B
    def plan_costs(self, x0: np.ndarray, knots: np.ndarray, basis: np.ndarray, cost_params: np.ndarray,
                   want_cost_matrix: bool = False) -> tuple[np.ndarray, np.ndarray | None]:
        x0, knots, basis, cost_params = _c(x0), _c(knots), _c(basis), _c(cost_params)
        N, K, _ = knots.shape
        H = basis.shape[0]
        assert basis.shape == (H, K) and cost_params.size == self.n_cost_params and x0.shape == (self.nq + self.nv,)
        reward = np.empty(N)
        cost = np.empty((N, H), dtype=np.float32) if want_cost_matrix else None
        cp = None if cost is None else cost.ctypes.data
        self._check(self._lib.b200mpc_plan_costs(self._h, _p(x0), _p(knots), N, K, _p(basis), H, _p(cost_params), cp, _p(reward)))
        return reward, cost

    def reward(self, states: np.ndarray, controls: np.ndarray, cost_params: np.ndarray, sensors: np.ndarray | None = None) -> np.ndarray:
        states, cost_params = _c(states), _c(cost_params)
        N, H, _ = states.shape
        # rewards that ignore the controls are called with controls=None by the reference's own tests
        controls = np.zeros((N, H, self.nu)) if controls is None else _c(controls)
        sensors = None if sensors is None else _c(sensors)
        assert controls.shape == (N, H, self.nu) and (sensors is None or sensors.shape == (N, H, self.nsensordata))
        out = np.empty(N)
        self._check(self._lib.b200mpc_reward_sensors(self._h, _p(states), _p(sensors), _p(controls), N, H, _p(cost_params), _p(out)))
        return out

    # ---- optimizer updates
    def update_mppi(self, knots: np.ndarray, rewards: np.ndarray, temperature: float) -> np.ndarray:
        knots, rewards = _c(knots), _c(rewards)
        N, K, nu = knots.shape
        out = np.empty((K, nu))
        self._check(self._lib.b200mpc_update_mppi(self._h, _p(knots), _p(rewards), N, K, float(temperature), _p(out)))
        return out

    def update_cem(self, knots: np.ndarray, rewards: np.ndarray, num_elites: int, sigma_min: float, sigma_max: float) -> tuple[np.ndarray, np.ndarray]:
        knots, rewards = _c(knots), _c(rewards)
        N, K, nu = knots.shape
        nom, sig = np.empty((K, nu)), np.empty((K, nu))
        self._check(self._lib.b200mpc_update_cem(self._h, _p(knots), _p(rewards), N, K, int(num_elites), float(sigma_min), float(sigma_max), _p(nom), _p(sig)))
        return nom, sig

    def update_ps(self, knots: np.ndarray, rewards: np.ndarray) -> np.ndarray:
        knots, rewards = _c(knots), _c(rewards)
        N, K, nu = knots.shape
        out = np.empty((K, nu))
        self._check(self._lib.b200mpc_update_ps(self._h, _p(knots), _p(rewards), N, K, _p(out)))
        return out

    # ---- fused plan step
    def plan_step(self, x0: np.ndarray, knots: np.ndarray, basis: np.ndarray, cost_params: np.ndarray, optimizer: str,
                  opt_params: np.ndarray, want_rewards: bool = True, n_elite: int = 0) -> dict:
        x0, knots, basis, cost_params = _c(x0), _c(knots), _c(basis), _c(cost_params)
        op = _c(np.atleast_1d(opt_params)) if opt_params is not None and np.size(opt_params) else np.zeros(1)
        N, K, nu = knots.shape
        H = basis.shape[0]
        assert basis.shape == (H, K) and cost_params.size == self.n_cost_params and x0.shape == (self.nq + self.nv,) and nu == self.nu
        nominal = np.empty((K, nu))
        sigma = np.empty((K, nu)) if optimizer == "cem" else None
        rewards = np.empty(N) if want_rewards else None
        elite = np.empty(max(n_elite, 1), dtype=np.int32)
        self._check(self._lib.b200mpc_plan_step(self._h, _p(x0), _p(knots), N, K, _p(basis), H, _p(cost_params), OPT_IDS[optimizer], _p(op),
                                                _p(nominal), _p(sigma), _p(rewards), elite.ctypes.data, int(n_elite)))
        return dict(nominal=nominal, sigma=sigma, rewards=rewards, elite=elite[:n_elite])

    # ---- Controller.update_action fast path (b200mpc_controller_step)
    def controller_step(self, rq: "_lib.StepRequest", stream: LegacyStream, n_normals: int) -> None:
        """Run one optimisation iteration described by ``rq`` (pointers already set), drawing the candidates' normals from numpy's
        global legacy stream.  Splits into phase 1 / tail / phase 2 when an odd number of normals is left after the head values."""
        rq.mt_key, rq.mt_pos = stream.key_addr, stream.pos_addr
        spec = self.speculative_sampling and n_normals >= self.SPECULATE_MIN
        rq.speculate, rq.use_speculated = int(spec), 0
        holds = self._lib.b200mpc_controller_speculation
        if spec and not (n_normals & 1) and holds(self._h, stream.key_addr, stream.pos_addr, n_normals):
            head = []        # the whole block was drawn during the previous step's GPU time, from exactly this generator state
            rq.use_speculated = 1
        else:
            head = stream.head(n_normals)
            if spec and holds(self._h, stream.key_addr, stream.pos_addr, (n_normals - len(head)) & ~1):
                rq.use_speculated = 1
        rq.n_head = len(head)
        for i, z in enumerate(head):
            rq.head[i] = z
        if (n_normals - len(head)) & 1:
            rq.phase = 1
            self._check(self._lib.b200mpc_controller_step(self._h, ctypes.addressof(rq)))
            rq.tail, rq.has_tail, rq.phase = stream.tail(), 1, 2
        else:
            rq.has_tail, rq.phase = 0, 0
        self._check(self._lib.b200mpc_controller_step(self._h, ctypes.addressof(rq)))

    def last_candidates(self, N: int, K: int) -> np.ndarray:
        out = np.empty((N, K, self.nu))
        self._check(self._lib.b200mpc_last_candidates(self._h, out.ctypes.data, int(N), int(K)))
        return out

    # ---- fused plan step with on-device sampling (perf mode; see include/b200mpc.h)
    def plan_step_sampled(self, x0: np.ndarray, nominal: np.ndarray, sigma: np.ndarray, lo: np.ndarray, hi: np.ndarray, num_rollouts: int,
                          basis: np.ndarray, cost_params: np.ndarray, optimizer: str, opt_params: np.ndarray, seed: int, counter: int,
                          index_offset: int = 0, want_rewards: bool = True, n_elite: int = 0, want_knots: bool = False) -> dict:
        x0, nominal, basis, cost_params = _c(x0), _c(nominal), _c(basis), _c(cost_params)
        K, nu = nominal.shape
        sigma = _c(np.broadcast_to(np.asarray(sigma, dtype=np.float64), (K, nu)))
        lo, hi = _c(np.broadcast_to(lo, (nu,))), _c(np.broadcast_to(hi, (nu,)))
        op = _c(np.atleast_1d(opt_params)) if opt_params is not None and np.size(opt_params) else np.zeros(1)
        H, N = basis.shape[0], int(num_rollouts)
        assert basis.shape == (H, K) and cost_params.size == self.n_cost_params and x0.shape == (self.nq + self.nv,) and nu == self.nu
        nom_out = np.empty((K, nu))
        sig_out = np.empty((K, nu)) if optimizer == "cem" else None
        rewards = np.empty(N) if want_rewards else None
        elite = np.empty(max(n_elite, 1), dtype=np.int32)
        elite_knots = np.empty((max(n_elite, 1), K, nu))
        knots = np.empty((N, K, nu)) if want_knots else None
        self._check(self._lib.b200mpc_plan_step_sampled(
            self._h, _p(x0), _p(nominal), _p(sigma), _p(lo), _p(hi), N, K, _p(basis), H, _p(cost_params), OPT_IDS[optimizer], _p(op),
            ctypes.c_ulonglong(int(seed) & (2**64 - 1)), ctypes.c_ulonglong(int(counter) & (2**64 - 1)), int(index_offset), _p(nom_out),
            _p(sig_out), _p(rewards), elite.ctypes.data, int(n_elite), _p(elite_knots), _p(knots)))
        return dict(nominal=nom_out, sigma=sig_out, rewards=rewards, elite=elite[:n_elite], elite_knots=elite_knots[:n_elite], knots=knots)


class MultiEngine(Engine):
    """One engine over several GPUs of THIS process (b200mpc_group_*, include/b200mpc.h): rollouts sharded along N over ``devices``
    (rollout 0, the un-noised nominal, on the first), the optimizer update over all of them on the first device.  Same methods as Engine;
    entry points that have no sharded form (reward(), update_*, plan_costs, plan_step_sampled) run on the first device."""

    def __init__(self, task: str, num_rollouts: int, devices: "list[int]", consts: np.ndarray | None = None) -> None:
        self._lib = _lib.load()
        self.task = task
        c = _c(task_consts(task) if consts is None else consts)
        dv = np.ascontiguousarray(devices, dtype=np.int32)
        g = ctypes.c_void_p()
        if self._lib.b200mpc_group_create(ctypes.byref(g), TASK_IDS[task], _p(c), c.size, dv.ctypes.data, len(dv), int(num_rollouts)):
            raise RuntimeError("b200mpc_group_create: " + self._lib.b200mpc_group_last_error(None).decode())
        self._g = g
        self._h = ctypes.c_void_p(self._lib.b200mpc_group_handle(g, 0))
        d = _lib.Dims()
        self._check(self._lib.b200mpc_get_dims(self._h, ctypes.byref(d)))
        self.nq, self.nv, self.nu, self.nsensordata, self.n_cost_params = d.nq, d.nv, d.nu, d.nsensordata, d.n_cost_params
        self.devices, self.device = [int(x) for x in dv], int(dv[0])

    def close(self) -> None:
        if getattr(self, "_g", None):
            self._lib.b200mpc_group_destroy(self._g)
            self._g, self._h = None, None

    def _gcheck(self, rc: int) -> None:
        if rc:
            raise RuntimeError(self._lib.b200mpc_group_last_error(self._g).decode())

    @property
    def num_rollouts(self) -> int:
        return self._lib.b200mpc_group_num_rollouts(self._g)

    @property
    def launch_count(self) -> int:
        return int(self._lib.b200mpc_group_launch_count(self._g))

    @property
    def contact_overflows(self) -> int:
        n = int(self._lib.b200mpc_group_contact_overflows(self._g))
        if n < 0:
            raise RuntimeError(self._lib.b200mpc_group_last_error(self._g).decode())
        return n

    def update(self, num_rollouts: int) -> None:
        self._gcheck(self._lib.b200mpc_group_update(self._g, int(num_rollouts)))

    def set_trace_capture(self, enable: bool) -> None:
        self._gcheck(self._lib.b200mpc_group_set_trace_capture(self._g, int(enable)))

    def elite_traces(self, idx: np.ndarray, H: int) -> np.ndarray:
        idx = np.ascontiguousarray(idx, dtype=np.int32)
        out = np.empty((len(idx), H, self.trace_width))
        self._gcheck(self._lib.b200mpc_group_elite_traces(self._g, idx.ctypes.data, len(idx), int(H), out.ctypes.data))
        return out

    def rollout(self, x0: np.ndarray, controls: np.ndarray, want_sensors: bool = True) -> tuple[np.ndarray, np.ndarray | None]:
        x0, controls = _c(x0), _c(controls)
        N, H, _ = controls.shape
        states = np.empty((N, H, self.nq + self.nv))
        sensors = np.empty((N, H, self.nsensordata)) if want_sensors else None
        self._gcheck(self._lib.b200mpc_group_rollout(self._g, _p(x0), int(x0.ndim == 2), _p(controls), N, H, _p(states), _p(sensors)))
        return states, sensors

    def plan_step(self, x0: np.ndarray, knots: np.ndarray, basis: np.ndarray, cost_params: np.ndarray, optimizer: str,
                  opt_params: np.ndarray, want_rewards: bool = True, n_elite: int = 0) -> dict:
        x0, knots, basis, cost_params = _c(x0), _c(knots), _c(basis), _c(cost_params)
        op = _c(np.atleast_1d(opt_params)) if opt_params is not None and np.size(opt_params) else np.zeros(1)
        N, K, nu = knots.shape
        H = basis.shape[0]
        assert basis.shape == (H, K) and cost_params.size == self.n_cost_params and x0.shape == (self.nq + self.nv,) and nu == self.nu
        nominal = np.empty((K, nu))
        sigma = np.empty((K, nu)) if optimizer == "cem" else None
        rewards = np.empty(N) if want_rewards else None
        elite = np.empty(max(n_elite, 1), dtype=np.int32)
        self._gcheck(self._lib.b200mpc_group_plan_step(self._g, _p(x0), _p(knots), N, K, _p(basis), H, _p(cost_params), OPT_IDS[optimizer], _p(op),
                                                       _p(nominal), _p(sigma), _p(rewards), elite.ctypes.data, int(n_elite)))
        return dict(nominal=nominal, sigma=sigma, rewards=rewards, elite=elite[:n_elite])

    def controller_step(self, rq: "_lib.StepRequest", stream: LegacyStream, n_normals: int) -> None:
        rq.mt_key, rq.mt_pos = stream.key_addr, stream.pos_addr
        spec = self.speculative_sampling and n_normals >= self.SPECULATE_MIN
        rq.speculate, rq.use_speculated = int(spec), 0
        holds = self._lib.b200mpc_group_controller_speculation
        if spec and not (n_normals & 1) and holds(self._g, stream.key_addr, stream.pos_addr, n_normals):
            head = []
            rq.use_speculated = 1
        else:
            head = stream.head(n_normals)
            if spec and holds(self._g, stream.key_addr, stream.pos_addr, (n_normals - len(head)) & ~1):
                rq.use_speculated = 1
        rq.n_head = len(head)
        for i, z in enumerate(head):
            rq.head[i] = z
        if (n_normals - len(head)) & 1:
            rq.phase = 1
            self._gcheck(self._lib.b200mpc_group_controller_step(self._g, ctypes.addressof(rq)))
            rq.tail, rq.has_tail, rq.phase = stream.tail(), 1, 2
        else:
            rq.has_tail, rq.phase = 0, 0
        self._gcheck(self._lib.b200mpc_group_controller_step(self._g, ctypes.addressof(rq)))

    def last_candidates(self, N: int, K: int) -> np.ndarray:
        out = np.empty((N, K, self.nu))
        self._gcheck(self._lib.b200mpc_group_last_candidates(self._g, out.ctypes.data, int(N), int(K)))
        return out
