"""CPU, world_size 2, gloo: the host-side logic of the N>1 path — sharding, the partial layout, the all_gather plumbing
(judo_b200.dist.gather_partials, the very function the NCCL path calls) and the combine algebra: sharded == unsharded."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank: int, world: int, port: int, q) -> None:  # noqa: ANN001
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist

    from judo_b200.dist import gather_partials, partial_width, shard_range
    from oracle import plan as op

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(123)  # same stream on every rank: the full problem is known everywhere, as in bench.py
        N, K, nu = 101, 4, 3              # odd N: uneven shards
        knots = rng.normal(size=(N, K, nu))
        rewards = -np.abs(rng.normal(size=N)) * 10
        rewards[7] = rewards[90] = rewards.max() + 1.0  # a tie across the two shards
        lo, hi = shard_range(N, world, rank)
        out = {}
        # MPPI
        p = torch.from_numpy(op.mppi_partial(knots[lo:hi], rewards[lo:hi], 0.05))
        assert p.numel() == partial_width("mppi", K * nu)
        allp = gather_partials(p, world).numpy()
        out["mppi"] = op.mppi_combine(allp, 0.05).reshape(K, nu)
        # CEM (k=3, ties: higher index first) and PS (k=1, ties: lower index first)
        for name, k, hi_first in (("cem", 3, True), ("ps", 1, False)):
            p = torch.from_numpy(op.topk_partial(knots[lo:hi], rewards[lo:hi], k, lo, hi_first))
            assert p.numel() == partial_width(name, K * nu, k)
            allp = gather_partials(p, world).numpy()
            elite, idx = op.topk_combine(allp, k, K * nu, hi_first)
            out[name] = (elite.reshape(-1, K, nu), idx)
        q.put((rank, lo, hi, out))
    finally:
        dist.destroy_process_group()


def test_shard_range_covers_everything():
    from judo_b200.dist import shard_range

    for n, w in ((8192, 8), (101, 2), (5, 8), (1024, 1)):
        spans = [shard_range(n, w, r) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1
    assert shard_range(8192, 8, 0) == (0, 1024)  # rollout 0 (the un-noised nominal) lives on rank 0


@pytest.mark.timeout(120)
def test_two_rank_gloo_update_equals_unsharded():
    import torch.multiprocessing as mp

    from oracle import plan as op

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=90) for _ in procs]
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    rng = np.random.default_rng(123)
    N, K, nu = 101, 4, 3
    knots = rng.normal(size=(N, K, nu))
    rewards = -np.abs(rng.normal(size=N)) * 10
    rewards[7] = rewards[90] = rewards.max() + 1.0
    ref_mppi = op.mppi_update(knots, rewards, 0.05)
    ref_cem, _ = op.cem_update(knots, rewards, 3, 0.0, 1e9)
    for rank, lo, hi, out in results:
        np.testing.assert_allclose(out["mppi"], ref_mppi, rtol=1e-12, atol=1e-14)   # identical on every rank
        elite, idx = out["cem"]
        assert list(idx[:2]) == [90, 7]                                              # tie: higher index first
        np.testing.assert_allclose(elite.mean(0), ref_cem, rtol=1e-12, atol=1e-14)
        elite, idx = out["ps"]
        assert idx[0] == 7                                                           # argmax: first maximum
        np.testing.assert_array_equal(elite[0], op.ps_update(knots, rewards))
