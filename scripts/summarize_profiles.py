"""Turn ncu outputs into the small tracked summaries under profiles/.

  python scripts/summarize_profiles.py launches <launches.csv> <out_summary.txt> "<command line>"
  python scripts/summarize_profiles.py full <raw.csv (ncu -i rep --page raw --csv)> <out.json> [traffic.json]
"""
import collections
import csv
import json
import re
import sys

KEEP = ["Kernel Name", "launch__grid_size", "launch__block_size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__waves_per_multiprocessor",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio"]


def short(name):
    name = re.sub(r"\(.*", "", name).replace("void ", "").replace("after::", "")
    return name.strip()


def to_bytes(v, unit):
    v = float(v)
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)


def launches(path, out, cmd):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10 and r[0].isdigit()]
    agg = collections.OrderedDict()
    for r in rows:
        k = short(r[4])
        ns = float(r[-1])
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += ns / 1e3
    tot = sum(a[1] for a in agg.values())
    with open(out, "w") as fh:
        fh.write(f"# {cmd}\n# (cold-cache, serialised per-launch times: compare SHARES)\n")
        fh.write(f"launches {len(rows)} total_us {tot:.0f}\n")
        for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            fh.write(f"{k:64s} n={n:5d} total_us={us:10.1f} avg_us={us / n:8.2f} share={100 * us / tot:5.1f}%\n")
    print(open(out).read())


def full(path, out, traffic=None):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    recs = []
    for r in rows[2:]:
        d = {}
        for k in KEEP:
            if k in hdr:
                i = hdr.index(k)
                d[k] = f"{r[i]} {units[i]}".strip()
        d["Kernel Name"] = short(r[hdr.index("Kernel Name")])
        i_r, i_w = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
        d["dram_bytes_per_launch"] = to_bytes(r[i_r], units[i_r]) + to_bytes(r[i_w], units[i_w])
        recs.append(d)
    json.dump(recs, open(out, "w"), indent=1)
    if traffic:
        by = collections.defaultdict(list)
        for d in recs:
            by[d["Kernel Name"]].append(d["dram_bytes_per_launch"])
        t = {"source": "ncu --set full --clock-control none (ncu flushes caches per launch: upper bound on the in-graph traffic), " + out}
        for k, v in by.items():
            t[k + "_bytes_per_launch"] = sum(v) / len(v)
        json.dump(t, open(traffic, "w"), indent=1)
    for d in recs:
        print(d["Kernel Name"][:60], d.get("gpu__time_duration.sum"), f"{d['dram_bytes_per_launch'] / 1e6:.1f} MB",
              d.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"))


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else "")
    else:
        full(sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else None)
