"""Summarise ncu output into the small JSON files kept under profiles/.

    python tools/ncu_summary.py full  <report.ncu-rep> <out.json>    # from `ncu --set full`
    python tools/ncu_summary.py list  <launches.csv>   <out.json>    # from `--metrics gpu__time_duration.sum --csv`

`full` averages, per kernel class (raygen_extend / bounce / accumulate / other), the counters the
north-star names: DRAM bytes, L2 bytes, issue-slot utilisation, warp execution efficiency, pipe
utilisation, warp instructions, registers, stall reasons per issue. `list` gives each class's
share of the step's kernel time.
"""
import collections
import csv
import json
import subprocess
import sys

FULL = {
    "gpu__time_duration.sum": "time_us",
    "dram__bytes_read.sum": "dram_read_bytes",
    "dram__bytes_write.sum": "dram_write_bytes",
    "lts__t_bytes.sum": "l2_bytes",
    "smsp__inst_executed.sum": "warp_inst",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
    "smsp__thread_inst_executed_per_inst_executed.ratio": "threads_per_warp_inst",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active": "pipe_alu_pct",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active": "pipe_fma_pct",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active": "pipe_xu_pct",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active": "pipe_lsu_pct",
    "smsp__warps_active.avg.per_cycle_active": "warps_active_per_scheduler",
    "smsp__warps_eligible.avg.per_cycle_active": "warps_eligible_per_scheduler",
    "launch__registers_per_thread": "registers",
    "launch__grid_size": "grid",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio": "stall_long_scoreboard",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio": "stall_short_scoreboard",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio": "stall_wait",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio": "stall_math_pipe_throttle",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio": "stall_not_selected",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio": "stall_branch_resolving",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active": "pipe_fp64_pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct_of_peak",
    "lts__t_sector_hit_rate.pct": "l2_hit_rate_pct",
}
UNIT_SCALE = {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1.0, "ms": 1e3, "us": 1.0, "ns": 1e-3, "msecond": 1e3,
              "usecond": 1.0, "nsecond": 1e-3, "second": 1e6, "s": 1e6}


def klass(name):
    for k in ("raygen_extend", "trace_kernel", "bounce", "accumulate", "resolve", "untile", "primary", "ref_visibility", "ref_shade"):
        if k in name:
            return k
    return "other"


def num(v):
    try:
        return float(v.replace(",", ""))
    except ValueError:
        return None


def full(rep, out):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    head, units = rows[0], rows[1]
    acc = collections.defaultdict(lambda: collections.defaultdict(list))
    names = collections.defaultdict(set)
    for r in rows[2:]:
        d = dict(zip(head, r))
        u = dict(zip(head, units))
        k = klass(d["Kernel Name"])
        names[k].add(d["Kernel Name"])
        for m, short in FULL.items():
            v = num(d.get(m, ""))
            if v is None:
                continue
            acc[k][short].append(v * UNIT_SCALE.get(u.get(m, ""), 1.0))
    res = {"_source": "ncu --set full --clock-control none (per-launch averages; cold-cache, serialised replays)", "report": rep}
    for k, ms in acc.items():
        res[k] = {"launches": len(ms["time_us"]), "kernels": sorted(names[k])}
        for short, vals in ms.items():
            res[k][short] = sum(vals) / len(vals)
        res[k]["dram_bytes_per_launch"] = res[k].get("dram_read_bytes", 0) + res[k].get("dram_write_bytes", 0)
    json.dump(res, open(out, "w"), indent=1)
    print(json.dumps({k: {s: round(v, 2) for s, v in d.items() if isinstance(v, float)} for k, d in res.items() if isinstance(d, dict)}, indent=1))


def launch_list(path, out):
    rows = list(csv.reader(open(path)))
    hi = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    idx = {h: i for i, h in enumerate(rows[hi])}
    tot = collections.Counter()
    cnt = collections.Counter()
    for r in rows[hi + 1:]:
        if len(r) <= idx["Metric Value"] or r[idx["Metric Name"]] != "gpu__time_duration.sum":
            continue
        v = num(r[idx["Metric Value"]])
        scale = UNIT_SCALE.get(r[idx["Metric Unit"]], 1.0)
        k = klass(r[idx["Kernel Name"]])
        tot[k] += v * scale
        cnt[k] += 1
    s = sum(tot.values())
    res = {"_source": "ncu --metrics gpu__time_duration.sum --clock-control none (launch list)", "file": path,
           "total_us": s, "classes": {k: {"launches": cnt[k], "us": tot[k], "share": tot[k] / s} for k in tot}}
    json.dump(res, open(out, "w"), indent=1)
    print(json.dumps(res["classes"], indent=1))


if __name__ == "__main__":
    {"full": full, "list": launch_list}[sys.argv[1]](sys.argv[2], sys.argv[3])
