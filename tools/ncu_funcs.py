#!/usr/bin/env python3
"""Attribute an `ncu --page source --csv` dump of one kernel to the device functions inside it.

usage: ncu_funcs.py <source.csv[.gz]> <nvdisasm -c output> <kernel mangled-name substring> [out.json]
The SASS offsets of the ncu rows (address - first address) are matched with the function labels nvdisasm prints
inside the kernel's .text section; per function: share of executed warp instructions, share of stall samples,
average active lanes, and the dominant stall reasons."""
import csv, gzip, io, json, re, sys


def main():
    src, dis, kname = sys.argv[1], sys.argv[2], sys.argv[3]
    op = gzip.open if src.endswith(".gz") else open
    rows = list(csv.reader(io.TextIOWrapper(op(src, "rb"))))
    h = rows[1]
    data = rows[2:]
    ia, ii, it, isamp = h.index("Address"), h.index("Instructions Executed"), h.index("Thread Instructions Executed"), h.index("# Samples")
    stall_cols = [(i, n) for i, n in enumerate(h) if n.startswith("stall_") and "Not Issued" not in n]
    a0 = int(data[0][ia], 16)
    # function labels by offset
    labels = []
    insec = False
    cur = None
    first = None
    for line in open(dis):
        if line.lstrip().startswith(".section"):
            insec = (".text." in line) and (kname in line) and (first is None or first == line)
            if insec:
                first = line                              # one instantiation only: a looser substring would mix the offsets of several
            continue
        if not insec:
            continue
        m = re.match(r"^(\$?[_A-Za-z][^\s:]*):\s*$", line)
        if m and not m.group(1).startswith(".L"):
            cur = m.group(1)
            continue
        m = re.search(r"/\*([0-9a-f]{4,})\*/", line)
        if m and cur is not None:
            off = int(m.group(1), 16)
            if not labels or labels[-1][1] != cur:
                labels.append((off, cur))
    labels.sort()
    offs = [o for o, _ in labels]
    import bisect
    agg = {}
    tot_i = tot_s = 0
    for r in data:
        off = int(r[ia], 16) - a0
        k = bisect.bisect_right(offs, off) - 1
        name = labels[k][1] if k >= 0 else "?"
        short = name.split("$")[-1] if "$" in name else "(kernel body)"
        m = re.search(r"E(\d+)([a-zA-Z_]\w*)E", short)
        mm = re.match(r"_ZN?K?\d*\w*?(\d+)([A-Za-z_]\w*)", short)
        d = agg.setdefault(short, {"inst": 0, "thr": 0, "samples": 0, "stalls": {}, "n_sass": 0})
        d["inst"] += int(r[ii]); d["thr"] += int(r[it]); d["samples"] += int(r[isamp]); d["n_sass"] += 1
        for i, n in stall_cols:
            v = int(r[i]) if r[i].isdigit() else 0
            if v:
                d["stalls"][n] = d["stalls"].get(n, 0) + v
        tot_i += int(r[ii]); tot_s += int(r[isamp])
    detail = __import__("os").environ.get("NCU_FUNC_DETAIL")
    if detail:                                           # hottest SASS rows of the functions whose name contains $NCU_FUNC_DETAIL
        isrc = h.index("Source")
        hot = []
        for r in data:
            off = int(r[ia], 16) - a0
            k = bisect.bisect_right(offs, off) - 1
            name = labels[k][1] if k >= 0 else "?"
            if detail in name:
                st = sorted(((int(r[i]) if r[i].isdigit() else 0, n[6:]) for i, n in stall_cols), reverse=True)[:2]
                hot.append((int(r[isamp]), off, int(r[ii]), r[isrc][:70], st))
        tot = sum(x[0] for x in hot)
        print("== %s: %d samples (%.1f%% of kernel), rows by offset with >= 0.5%% of the function's samples" % (detail, tot, 100.0 * tot / max(tot_s, 1)))
        for smp, off, ninst, text, st in hot:
            if smp >= 0.005 * tot:
                print("  %06x  inst %10d  samples %7d (%4.1f%%)  %-70s %s" % (off, ninst, smp, 100.0 * smp / max(tot, 1), text, st))
    out = []
    for name, d in agg.items():
        st = sorted(d["stalls"].items(), key=lambda kv: -kv[1])[:4]
        out.append({"function": name, "sass_instructions": d["n_sass"], "inst_pct": round(100.0 * d["inst"] / max(tot_i, 1), 2),
                    "stall_sample_pct": round(100.0 * d["samples"] / max(tot_s, 1), 2),
                    "active_lanes": round(d["thr"] / max(d["inst"], 1), 1),
                    "top_stalls_pct_of_function": {k[6:]: round(100.0 * v / max(d["samples"], 1), 1) for k, v in st}})
    out.sort(key=lambda x: -x["stall_sample_pct"])
    res = {"kernel": kname, "total_warp_instructions": tot_i, "total_stall_samples": tot_s, "functions": out}
    if len(sys.argv) > 4:
        json.dump(res, open(sys.argv[4], "w"), indent=1)
    for o in out[:45]:
        print("%-60s sass %5d inst %5.1f%% samples %5.1f%% lanes %4.1f  %s" % (o["function"][:60], o["sass_instructions"], o["inst_pct"], o["stall_sample_pct"],
                                                                            o["active_lanes"], o["top_stalls_pct_of_function"]))


if __name__ == "__main__":
    main()
