"""gpurun_out/r02_counts_<workload>.csv (tools/r02_ncu_counts.sh) -> profiles/r02_counts.json, read by bench.py for roofline.traffic and the
fp64-issue roofline.  Counts are per launch of the dominant kernel (rollout + cost, no epilogue) at the bench's own size."""
import csv, glob, json, os, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out = {"_comment": "per-launch counts of the dominant kernel from `ncu --metrics ... -k regex:<kernel> -s 6 -c 1 python bench.py --workload W "
                   "--steps 3 --warmup 3 --no-extras` on one B200 (tools/r02_ncu_counts.sh); times under the profiler are NOT used"}
for f in sorted(glob.glob(os.path.join(ROOT, "gpurun_out", "r02_counts_*.csv"))):
    name = os.path.basename(f)[len("r02_counts_"):-4]
    rows = [r for r in csv.reader(open(f)) if len(r) > 5]
    if not rows:
        continue
    hdr = rows[0]
    im, iv, iu = hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    m = {}
    for r in rows[1:]:
        v = float(r[iv].replace(",", ""))
        u = r[iu]
        if u in ("Mbyte", "MByte"): v *= 1e6
        if u in ("Kbyte", "KByte"): v *= 1e3
        if u in ("Gbyte", "GByte"): v *= 1e9
        if u == "ms": v *= 1e-3
        if u == "us": v *= 1e-6
        if u in ("ns", "nsecond"): v *= 1e-9
        m[r[im]] = v
        kern = r[hdr.index("Kernel Name")]
    wi, ti = m["smsp__inst_executed.sum"], m["smsp__thread_inst_executed.sum"]
    cyc = m["sm__cycles_elapsed.max"]
    out[name] = {"kernel": kern[:80], "dram_bytes": int(m["dram__bytes_read.sum"] + m["dram__bytes_write.sum"]),
                 "dram_bytes_read": int(m["dram__bytes_read.sum"]), "dram_bytes_write": int(m["dram__bytes_write.sum"]),
                 "warp_inst": int(wi), "thread_inst": int(ti), "active_lanes": round(ti / wi, 2),
                 "fp64_warp_inst": int(m["sm__inst_executed_pipe_fp64.sum"]), "fp64_pipe_cycles_active_avg_per_sm": m["sm__pipe_fp64_cycles_active.avg"],
                 "sm_cycles_elapsed_max": int(cyc), "ipc_per_sm": round(wi / cyc / 148, 3),
                 "warps_active_per_sm": round(m.get("sm__warps_active.avg.per_cycle_active", 0), 2),
                 "issue_active_pct": round(m.get("smsp__issue_active.avg.pct_of_peak_sustained_active", 0), 2),
                 "profiled_duration_s": m["gpu__time_duration.sum"]}
json.dump(out, open(os.path.join(ROOT, "profiles", "r02_counts.json"), "w"), indent=1)
print(json.dumps(out, indent=1))
