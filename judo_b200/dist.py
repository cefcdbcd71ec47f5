"""Sharded, device-resident plan step: one process per GPU, rollouts split along N.

Every rank rolls out its own slice with the fused kernel, whose epilogue reduces it to ONE partial
(MPPI: [beta, S, V[K*nu]]; CEM/PS: k x [reward, global index, knots[K*nu]]), the partials are exchanged with a single
``all_gather`` (NCCL over NVLink; a few hundred bytes per rank — latency-bound) and every rank runs the same tiny
combine kernel, so all ranks hold the identical nominal knots.  world_size == 1 skips the collective.

torch is plumbing here: it owns the device allocations, the stream and the process group; all compute goes through
the C ABI's resident (``*_dev``) entry points.
"""

from __future__ import annotations

import ctypes

import numpy as np


def shard_range(n_total: int, world_size: int, rank: int) -> tuple[int, int]:
    """Contiguous slice [lo, hi) of the global rollout axis owned by ``rank`` (rollout 0, the un-noised nominal,
    lives on rank 0).  Remainders go to the lowest ranks."""
    base, rem = divmod(n_total, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def partial_width(optimizer: str, knu: int, k: int = 1) -> int:
    """Doubles per rank in the exchanged partial."""
    return 2 + knu if optimizer == "mppi" else k * (2 + knu)


def gather_partials(local, world_size: int, group=None):  # noqa: ANN001
    """all_gather of one fixed-size partial per rank -> (world_size, width) tensor.  Works for NCCL (cuda tensors)
    and gloo (cpu tensors); used as is by the CPU tests of the N>1 path."""
    import torch
    import torch.distributed as dist

    if world_size == 1:
        return local.reshape(1, -1)
    flat = local.reshape(-1).contiguous()
    out = torch.empty(world_size * flat.numel(), dtype=flat.dtype, device=flat.device)
    dist.all_gather_into_tensor(out, flat, group=group)
    return out.view(world_size, flat.numel())


class ShardedPlanner:
    """Resident plan step for one rank.  Inputs are uploaded once (``set_problem`` / ``set_knots``); ``step`` launches
    rollout+cost -> partial -> [all_gather] -> combine on the current torch stream and returns device tensors."""

    def __init__(self, task: str, n_local: int, device: int = 0, rank: int = 0, world_size: int = 1, group=None) -> None:  # noqa: ANN001
        import torch

        from judo_b200 import _lib
        from judo_b200.engine import Engine

        self.torch = torch
        self.lib = _lib.load()
        self.engine = Engine(task, n_local, device=device)
        self.dev = torch.device("cuda", device)
        self.rank, self.world_size, self.group = rank, world_size, group
        self.n_local = n_local
        self.peer_exchange = False
        self.nu = self.engine.nu
        self.nx = self.engine.nq + self.engine.nv

    def _check(self, rc: int) -> None:
        if rc:
            raise RuntimeError(self.lib.b200mpc_last_error(self.engine.handle).decode())

    def enable_peer_exchange(self) -> bool:
        """Open every rank's exchange buffer through CUDA IPC so that the MPPI update's cross-GPU step happens INSIDE the rollout
        kernel (P2P stores over NVLink + flags) instead of an NCCL all_gather + a combine launch (MPPI, CEM and PS on the fused kernels)."""
        import torch.distributed as dist

        t = self.torch
        mine = np.zeros(64, dtype=np.uint8)
        ok = self.lib.b200mpc_exchange_create(self.engine.handle, self.world_size, self.rank, mine.ctypes.data) == 0
        if self.world_size > 1:
            buf = t.from_numpy(mine).to(self.dev)
            allh = t.empty(self.world_size * 64, dtype=t.uint8, device=self.dev)
            dist.all_gather_into_tensor(allh, buf, group=self.group)
            handles = np.ascontiguousarray(allh.cpu().numpy())
        else:
            handles = mine
        ok = ok and self.lib.b200mpc_exchange_open(self.engine.handle, handles.ctypes.data) == 0
        if self.world_size > 1:
            # every rank must take the same path: one failed IPC mapping anywhere -> everybody stays on the all_gather path
            flag = t.tensor([1 if ok else 0], dtype=t.int32, device=self.dev)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=self.group)
            ok = bool(flag.item())
        self.peer_exchange = ok
        return ok

    @staticmethod
    def wire_local(planners: list) -> None:
        """Peer exchange between planners that live in ONE process (one per stream / device): buffers are exchanged as plain device pointers.
        Run one step per planner BEFORE wiring (it sizes the handle's scratch buffers): an allocation inside a step may wait for the
        device while a peer's kernel is already spinning on this rank's partial."""
        world = len(planners)
        bufs = (ctypes.c_void_p * world)()
        for r, p in enumerate(planners):
            mine = np.zeros(64, dtype=np.uint8)
            p._check(p.lib.b200mpc_exchange_create(p.engine.handle, world, r, mine.ctypes.data))
            b = ctypes.c_void_p()
            p._check(p.lib.b200mpc_exchange_buffer(p.engine.handle, ctypes.byref(b)))
            bufs[r] = b.value
        for p in planners:
            p._check(p.lib.b200mpc_exchange_open_local(p.engine.handle, bufs))
            p.peer_exchange = True

    def align(self) -> None:
        """Line the ranks' GPUs up on the current stream (no data moves): a one-warp kernel that signals every peer and waits for all."""
        st = ctypes.c_void_p(self.torch.cuda.current_stream(self.dev).cuda_stream)
        self._check(self.lib.b200mpc_exchange_align_dev(self.engine.handle, st))

    def exchange_stamps(self) -> tuple[int, int, int]:
        """%globaltimer (ns) of the last in-kernel-exchange step: kernel entry, partial published, all peers seen."""
        out = (ctypes.c_ulonglong * 3)()
        self._check(self.lib.b200mpc_exchange_stamps(self.engine.handle, out))
        return int(out[0]), int(out[1]), int(out[2])

    def exchange_gap_stamps(self) -> tuple[int, int]:
        """%globaltimer (ns): exit of the last line-up kernel, last instruction of the last in-kernel exchange."""
        out = (ctypes.c_ulonglong * 2)()
        self._check(self.lib.b200mpc_exchange_align_stamp(self.engine.handle, out))
        return int(out[0]), int(out[1])

    def set_problem(self, x0: np.ndarray, basis: np.ndarray, cost_params: np.ndarray, want_cost_matrix: bool = True) -> None:
        t = self.torch
        self.H, self.K = basis.shape
        self.knu = self.K * self.nu
        self.d_x0 = t.as_tensor(np.ascontiguousarray(x0, dtype=np.float64), device=self.dev)
        self.d_basis = t.as_tensor(np.ascontiguousarray(basis, dtype=np.float64), device=self.dev)
        self.d_params = t.as_tensor(np.ascontiguousarray(cost_params, dtype=np.float64), device=self.dev)
        self.d_reward = t.empty(self.n_local, dtype=t.float64, device=self.dev)
        self.d_cost = t.empty((self.n_local, self.H), dtype=t.float32, device=self.dev) if want_cost_matrix else None
        self.d_nominal = t.empty(self.knu, dtype=t.float64, device=self.dev)
        self.d_sigma = t.empty(self.knu, dtype=t.float64, device=self.dev)
        self.d_elite = t.empty(64, dtype=t.float64, device=self.dev)

    def set_knots(self, knots_local: np.ndarray) -> None:
        assert knots_local.shape == (self.n_local, self.K, self.nu)
        self.d_knots = self.torch.as_tensor(np.ascontiguousarray(knots_local, dtype=np.float64), device=self.dev)

    def step(self, optimizer: str, opt_params: np.ndarray, index_offset: int = 0, n_elite: int = 0):  # noqa: ANN201
        """One resident plan step; returns the (K*nu,) nominal device tensor (identical on every rank).

        world_size == 1: ONE launch (rollout + cost + update fused).  world_size > 1: the fused kernel leaves this rank's
        partial, one all_gather moves world_size partials, one tiny combine kernel finishes."""
        t = self.torch
        from judo_b200.engine import OPT_IDS

        st = ctypes.c_void_p(t.cuda.current_stream(self.dev).cuda_stream)
        h = self.engine.handle
        P = lambda x: ctypes.c_void_p(0 if x is None else x.data_ptr())  # noqa: E731
        op = np.ascontiguousarray(np.atleast_1d(opt_params), dtype=np.float64) if np.size(opt_params) else np.zeros(1)
        opp = op.ctypes.data
        single = self.world_size == 1
        k_x = int(op[0]) if optimizer == "cem" else 1
        fits = (2 + self.knu if optimizer == "mppi" else k_x * (2 + self.knu)) <= 98 and k_x <= 8  # EP_XCHG_STRIDE doubles per rank slot, EP_MAXK elites
        if (not single or getattr(self, "force_peer", False)) and self.peer_exchange and fits and self.engine.task not in ("leap_cube", "leap_cube_down", "fr3_pick"):
            # ONE kernel per rank: rollout + cost + P2P exchange of the partials + final update
            self._check(self.lib.b200mpc_plan_step_dev(h, P(self.d_x0), P(self.d_knots), self.n_local, self.K, P(self.d_basis), self.H,
                                                       P(self.d_params), OPT_IDS[optimizer], opp, 2, int(index_offset), 0, P(self.d_cost),
                                                       P(self.d_reward), P(self.d_nominal), P(self.d_sigma), P(self.d_elite), P(None), st))
            return self.d_nominal
        k = int(op[0]) if optimizer == "cem" else 1
        width = 2 + self.knu if optimizer == "mppi" else k * (2 + self.knu)
        part = None if single else t.empty(width, dtype=t.float64, device=self.dev)
        self._check(self.lib.b200mpc_plan_step_dev(h, P(self.d_x0), P(self.d_knots), self.n_local, self.K, P(self.d_basis), self.H,
                                                   P(self.d_params), OPT_IDS[optimizer], opp, int(single), int(index_offset),
                                                   int(n_elite) if single else 0, P(self.d_cost), P(self.d_reward), P(self.d_nominal),
                                                   P(self.d_sigma), P(self.d_elite), P(part), st))
        if not single:
            allp = gather_partials(part, self.world_size, self.group)
            if optimizer == "mppi":
                self._check(self.lib.b200mpc_mppi_combine_dev(h, P(allp), self.world_size, self.knu, float(op[0]), P(self.d_nominal), st))
            else:
                hi = 1 if optimizer == "cem" else 0
                smin, smax = (float(op[1]), float(op[2])) if optimizer == "cem" else (0.0, 0.0)
                self._check(self.lib.b200mpc_topk_combine_dev(h, P(allp), self.world_size, self.knu, k, hi, smin, smax, P(self.d_nominal),
                                                              P(self.d_sigma) if optimizer == "cem" else ctypes.c_void_p(0), P(self.d_elite), st))
            self._keep = (part, allp)  # keep the buffers alive until the stream has consumed them
        return self.d_nominal

