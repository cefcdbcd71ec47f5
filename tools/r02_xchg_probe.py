"""2-rank probe (torchrun): where do the ~10 us between the fused kernel's own entry->done time and the event-timed step go?
Per step, device-side %globaltimer stamps: line-up kernel exit, rollout kernel entry, partial published, all peers seen; CUDA events around
the step.  Also the SAME aligned loop with the single-GPU fused step (finalize=1, no exchange) on every rank."""
import ctypes, json, os, statistics, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from judo_b200.dist import ShardedPlanner, shard_range

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
dev = torch.device("cuda", lr)
w = dict(bench.WORKLOADS["cartpole_mppi"])
n_local = w["n_rollouts"]
task, opt, x0, knots_all, basis, params, _ = bench.problem(w, n_local * world)
lo, hi = shard_range(n_local * world, world, rank)
pl = ShardedPlanner("cartpole", n_local, device=lr, rank=rank, world_size=world)
assert pl.enable_peer_exchange()
pl.set_problem(x0, basis, params, want_cost_matrix=True)
pl.set_knots(np.ascontiguousarray(knots_all[lo:hi]))
op = opt.fused_params()
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

def run(mode, flush_on, n=60):
    ev, gt = [], []
    for i in range(n):
        if flush_on: flush.zero_()
        pl.align()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        if mode == "xchg":
            pl.step("mppi", op, index_offset=lo)
        else:
            pl.world_size = 1; pl.step("mppi", op, index_offset=lo); pl.world_size = world
        e.record()
        torch.cuda.synchronize(dev)
        ev.append(s.elapsed_time(e) * 1e3)
        a = pl.exchange_gap_stamps()
        if mode == "xchg":
            t_in, t_pub, t_done = pl.exchange_stamps()
            gt.append(((t_in - a[0]) * 1e-3, (t_pub - t_in) * 1e-3, (t_done - t_pub) * 1e-3, (a[1] - t_done) * 1e-3))
    ev = ev[10:]; gt = gt[10:]
    out = {"event_us_mean": statistics.mean(ev), "event_us_p50": statistics.median(ev)}
    if gt:
        g = np.array(gt)
        out.update(align_exit_to_entry_us=float(np.median(g[:, 0])), entry_to_published_us=float(np.median(g[:, 1])), published_to_all_seen_us=float(np.median(g[:, 2])), all_seen_to_last_instr_us=float(np.median(g[:, 3])))
    return out

def run_queued(mode, n=60):
    """as bench.py: everything queued, no host sync per step"""
    S = [torch.cuda.Event(enable_timing=True) for _ in range(n)]; E = [torch.cuda.Event(enable_timing=True) for _ in range(n)]
    for i in range(n):
        flush.zero_(); pl.align(); S[i].record()
        if mode == "xchg": pl.step("mppi", op, index_offset=lo)
        else:
            pl.world_size = 1; pl.step("mppi", op, index_offset=lo); pl.world_size = world
        E[i].record()
    torch.cuda.synchronize(dev)
    t = [s.elapsed_time(e) * 1e3 for s, e in zip(S, E)][10:]
    return {"event_us_mean": statistics.mean(t), "event_us_p50": statistics.median(t)}

res = {}
for name, f in (("xchg_sync_flush", lambda: run("xchg", True)), ("xchg_sync_noflush", lambda: run("xchg", False)), ("local_sync_flush", lambda: run("local", True)),
                ("xchg_queued", lambda: run_queued("xchg")), ("local_queued", lambda: run_queued("local"))):
    dist.barrier(); torch.cuda.synchronize(dev)
    res[name] = f()
allr = [None] * world
dist.all_gather_object(allr, res)
if rank == 0:
    print(json.dumps({"world": world, "per_rank": allr}))
dist.destroy_process_group()
