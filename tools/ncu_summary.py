#!/usr/bin/env python3
"""Summarise one kernel launch of an .ncu-rep (ncu --set full) into a small JSON for profiles/.

usage: ncu_summary.py <report.ncu-rep> <out.json> [note]
"""
import csv, json, subprocess, sys

KEYS = {
    "gpu__time_duration.sum": "time",
    "launch__registers_per_thread": "registers_per_thread",
    "launch__grid_size": "grid", "launch__block_size": "block",
    "smsp__inst_executed.sum": "warp_instructions",
    "smsp__thread_inst_executed_per_inst_executed.ratio": "active_lanes_per_instruction",
    "sm__inst_executed.avg.per_cycle_active": "ipc_per_sm",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
    "sm__icc_request_hit_rate.pct": "icache_hit_pct",
    "gcc__cache_requests_type_instruction.sum.pct_of_peak_sustained_elapsed": "gpc_instruction_cache_requests_pct_of_peak",
    "l1tex__t_sector_hit_rate.pct": "l1_hit_pct", "lts__t_sector_hit_rate.pct": "l2_hit_pct",
    "dram__bytes_read.sum": "dram_read", "dram__bytes_write.sum": "dram_write",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed": "dram_throughput_pct",
}


def main():
    rep, out = sys.argv[1], sys.argv[2]
    note = sys.argv[3] if len(sys.argv) > 3 else ""
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    h, u, v = rows[0], rows[1], rows[2]
    d = {"report": rep.split("/")[-1], "note": note, "kernel": v[h.index("Kernel Name")] if "Kernel Name" in h else ""}
    stalls = {}
    for i, n in enumerate(h):
        if n in KEYS:
            d[KEYS[n]] = "%s %s" % (v[i], u[i])
        if n.startswith("smsp__average_warps_issue_stalled_") and n.endswith("_per_issue_active.ratio") and "not_issued" not in n:
            try:
                x = float(v[i].replace(",", ""))
            except ValueError:
                continue
            if x >= 0.05:
                stalls[n[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]] = round(x, 2)
    d["stall_cycles_per_issued_instruction"] = dict(sorted(stalls.items(), key=lambda kv: -kv[1]))
    json.dump(d, open(out, "w"), indent=1)
    print(json.dumps(d, indent=1))


if __name__ == "__main__":
    main()
