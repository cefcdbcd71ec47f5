"""GPU: the multi-GPU plan step — in-kernel P2P exchange (CUDA IPC + NVLink) and the NCCL all_gather fallback."""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _problem(N):
    from judo_b200.spline import spline_basis

    rng = np.random.default_rng(21)
    K, H = 4, 40
    x0 = np.array([0.3, 2.9, 0.1, -0.2])
    knots = np.clip(0.4 * rng.normal(size=(N, K, 1)), -1.8, 1.8)
    basis = spline_basis(np.linspace(0, 1.6, K), 0.04 * np.arange(H), "zero")
    params = np.array([10, 10, 0.1, 0.1, 0.01, 0.1])
    return x0, knots, basis, params


@pytest.mark.parametrize("optimizer,opt_params", [("mppi", [0.05]), ("cem", [3, 0.1, 1.0]), ("ps", [])])
def test_peer_exchange_path_on_one_rank(optimizer, opt_params):
    """finalize=2 with world_size 1: the exchange buffer, flags and epochs work (several consecutive steps)."""
    from judo_b200.dist import ShardedPlanner
    from oracle import plan as op

    N = 512
    x0, knots, basis, params = _problem(N)
    pl = ShardedPlanner("cartpole", N)
    pl.enable_peer_exchange()
    pl.force_peer = True
    pl.set_problem(x0, basis, params)
    pl.set_knots(knots)
    for _ in range(4):
        nominal = pl.step(optimizer, np.array(opt_params)).cpu().numpy().reshape(4, 1)
    rewards = pl.d_reward.cpu().numpy()
    ref = {"mppi": lambda: op.mppi_update(knots, rewards, 0.05), "cem": lambda: op.cem_update(knots, rewards, 3, 0.1, 1.0)[0],
           "ps": lambda: op.ps_update(knots, rewards)}[optimizer]()
    np.testing.assert_allclose(nominal, ref, rtol=1e-11, atol=1e-13)
    if optimizer == "cem":
        np.testing.assert_allclose(pl.d_sigma.cpu().numpy().reshape(4, 1), op.cem_update(knots, rewards, 3, 0.1, 1.0)[1], rtol=1e-11)


def _rank_main(rank, world, port, mode, q, optimizer="mppi", opt_params=(0.05,)):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist

    from judo_b200.dist import ShardedPlanner, shard_range

    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        N = 1000  # uneven shards
        x0, knots, basis, params = _problem(N)
        lo, hi = shard_range(N, world, rank)
        pl = ShardedPlanner("cartpole", hi - lo, device=rank, rank=rank, world_size=world)
        if mode == "peer":
            pl.enable_peer_exchange()
        pl.set_problem(x0, basis, params)
        pl.set_knots(knots[lo:hi])
        outs = []
        for _ in range(3):
            outs.append(pl.step(optimizer, np.array(opt_params), index_offset=lo).cpu().numpy().copy())
        torch.cuda.synchronize()
        q.put((rank, outs, pl.d_reward.cpu().numpy(), lo, hi))
    except Exception as e:  # noqa: BLE001 — report instead of letting the parent wait for its time-out
        q.put((rank, repr(e), None, 0, 0))
        raise
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("optimizer,opt_params", [("mppi", (0.05,)), ("cem", (3, 0.1, 1.0)), ("ps", ())])
@pytest.mark.parametrize("mode", ["peer", "nccl"])
@pytest.mark.timeout(300)
def test_two_gpu_step_matches_unsharded(mode, optimizer, opt_params):
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp

    from oracle import plan as op

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + (os.getpid() * 7 + hash((mode, optimizer)) % 97) % 2000
    procs = [ctx.Process(target=_rank_main, args=(r, 2, port, mode, q, optimizer, opt_params)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for r in res:
        assert r[2] is not None, f"rank {r[0]} failed: {r[1]}"
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    x0, knots, basis, params = _problem(1000)
    rewards = np.empty(1000)
    for rank, outs, r, lo, hi in res:
        rewards[lo:hi] = r
    ref = {"mppi": lambda: op.mppi_update(knots, rewards, 0.05), "cem": lambda: op.cem_update(knots, rewards, 3, 0.1, 1.0)[0],
           "ps": lambda: op.ps_update(knots, rewards)}[optimizer]().ravel()
    for rank, outs, r, lo, hi in res:
        for o in outs:
            np.testing.assert_allclose(o, ref, rtol=1e-11, atol=1e-13)   # identical nominal on every rank, every step
