"""Per-kernel table of an `ncu --set full ... --page raw --csv` export (one row per kernel name: launches, total and longest
duration, and the key metrics of the LONGEST launch), plus the DRAM traffic per kernel family as JSON for bench.py.

    python tools/ncu_table.py gpurun_out/r2final_full_raw.csv [profiles/r2_traffic.json]
"""
import json
import re
import sys

import pandas as pd

path = sys.argv[1]
df = pd.read_csv(path, low_memory=False).iloc[1:].copy()
M = {"dur": "gpu__time_duration.sum", "grid": "launch__grid_size", "blk": "launch__block_size", "regs": "launch__registers_per_thread",
     "tensor": "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "issue": "smsp__issue_active.avg.pct_of_peak_sustained_active",
     "warps": "sm__warps_active.avg.pct_of_peak_sustained_active", "dram": "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
     "lts": "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1": "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
     "fma": "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "alu": "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
     "lsb": "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "inst": "smsp__inst_executed.sum",
     "rd": "dram__bytes_read.sum", "wr": "dram__bytes_write.sum", "l2hit": "lts__t_sector_hit_rate.pct"}
units = pd.read_csv(path, low_memory=False).iloc[0]
for c in M.values():
    df[c] = pd.to_numeric(df[c].astype(str).str.replace(",", ""), errors="coerce")


def short(n):
    n = re.sub(r"\(.*", "", n).replace("samble::", "").replace("void ", "")
    return n[:48]


def to_bytes(col):
    u = str(units[col]).lower()
    return {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)


df["k"] = df["Kernel Name"].map(short)
rows = []
for k, d in df.groupby("k"):
    rows.append((d[M["dur"]].sum(), k, len(d), d.loc[d[M["dur"]].idxmax()], d))
rows.sort(key=lambda r: -r[0])
print(f"# ncu --set full --clock-control none: {path}\n")
print("One row per kernel; metrics are those of its LONGEST launch (durations in us, cold-cache and serialised: compare shares).\n")
print("| kernel | launches | total us | longest us | grid x block | regs | tensor % | issue % | warps % | fma % | alu % | dram % | lts % | l1 % | L2 hit % | long-sb | Minst | DRAM MB (longest) |")
print("|---|--:|--:|--:|---|--:|--:|--:|--:|--:|--:|--:|--:|--:|--:|--:|--:|--:|")
fam = {}
for tot, k, n, t, d in rows:
    mb = (t[M["rd"]] * to_bytes(M["rd"]) + t[M["wr"]] * to_bytes(M["wr"])) / 1e6
    print(f"| `{k}` | {n} | {tot:.1f} | {t[M['dur']]:.1f} | {int(t[M['grid']])}x{int(t[M['blk']])} | {int(t[M['regs']])} | {t[M['tensor']]:.1f} | "
          f"{t[M['issue']]:.1f} | {t[M['warps']]:.1f} | {t[M['fma']]:.1f} | {t[M['alu']]:.1f} | {t[M['dram']]:.1f} | {t[M['lts']]:.1f} | {t[M['l1']]:.1f} | "
          f"{t[M['l2hit']]:.0f} | {t[M['lsb']]:.1f} | {t[M['inst']] / 1e6:.2f} | {mb:.1f} |")
    name = re.sub(r"<.*", "", k)
    f = fam.setdefault(name, {"dram_bytes_per_step": 0, "launches_per_step": 0, "us_per_step_cold": 0.0})
    f["dram_bytes_per_step"] += int((d[M["rd"]] * to_bytes(M["rd"]) + d[M["wr"]] * to_bytes(M["wr"])).sum())
    f["launches_per_step"] += int(n)
    f["us_per_step_cold"] += float(tot)
print(f"\ntotal of the captured launches: {df[M['dur']].sum():.1f} us")
if len(sys.argv) > 2:
    json.dump({"source": path, "families": fam}, open(sys.argv[2], "w"), indent=1)
