"""Turn ncu CSV exports into the small, reviewable summaries committed under profiles/.

  python tools/summarize_ncu.py launches <launches.csv> > profiles/<name>.md      # per-kernel time shares
  python tools/summarize_ncu.py raw <raw.csv> > profiles/<name>.md                 # key metrics of a --set full capture
"""
import collections
import csv
import sys

KEYS = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio"]


def launches(path):
    rows = list(csv.reader(open(path)))
    start = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    h = rows[start]
    ki, vi = h.index("Kernel Name"), h.index("Metric Value")
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows[start + 1:]:
        if len(r) <= vi:
            continue
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        name = r[ki].split("(")[0].replace("void ", "")
        agg[name][0] += 1
        agg[name][1] += v
    tot = sum(v[1] for v in agg.values())
    print(f"# ncu launch list: {path}\n\ngpu__time_duration.sum per kernel (cold-cache, serialised; compare SHARES). total {tot / 1e3:.1f} us\n")
    print("| share | total us | launches | kernel |\n|---:|---:|---:|---|")
    for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
        print(f"| {100 * t / tot:.1f}% | {t / 1e3:.1f} | {c} | `{n[:110]}` |")


def raw(path):
    rows = list(csv.reader(open(path)))
    h = rows[0]
    print(f"# ncu --set full: {path}\n")
    for r in rows[2:]:
        d = dict(zip(h, r))
        print(f"## `{d['Kernel Name'][:120]}`\n")
        print("| metric | value |\n|---|---:|")
        for k in KEYS:
            if k in d:
                print(f"| {k} | {d[k]} |")
        print()


def traffic(path):
    """JSON for bench.py's roofline.traffic: DRAM bytes (read + write) per kernel family, summed over the launches of
    ONE step captured with `ncu --set full` (units row of the raw page gives the scale of each column)."""
    import json
    import re

    rows = list(csv.reader(open(path)))
    h, units = rows[0], rows[1]
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    cols = [(i, scale[units[i]]) for i, k in enumerate(h) if k in ("dram__bytes_read.sum", "dram__bytes_write.sum")]
    ki = h.index("Kernel Name")
    out = collections.defaultdict(float)
    n = collections.defaultdict(int)
    for r in rows[2:]:
        m = re.search(r"(\w+_kernel)", r[ki])
        name = m.group(1) if m else r[ki].split("(")[0]
        out[name] += sum(float(r[i].replace(",", "")) * sc for i, sc in cols)
        n[name] += 1
    print(json.dumps({"source": path, "dram_bytes_per_step": {k: round(v) for k, v in out.items()},
                      "launches_per_step": dict(n)}, indent=1))


if __name__ == "__main__":
    {"launches": launches, "raw": raw, "traffic": traffic}[sys.argv[1]](sys.argv[2])
