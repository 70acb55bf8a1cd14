"""Summarise an ncu report of the solve kernel: headline metrics, stall reasons, and stall samples per
source function / line of ipddp_solver.h (needs the kernel built with -lineinfo and --import-source on).
    python tools/ncu_summary.py gpurun_out/prof.ncu-rep [--lines 40]
"""
import bisect
import collections
import csv
import io
import re
import subprocess
import sys

rep = sys.argv[1]
nlines = int(sys.argv[sys.argv.index("--lines") + 1]) if "--lines" in sys.argv else 30


def ncu(*args):
    return subprocess.run(["ncu", "-i", rep] + list(args), capture_output=True, text=True).stdout


raw = list(csv.reader(io.StringIO(ncu("--page", "raw", "--csv"))))
hdr, units, vals = raw[0], raw[1], raw[2]
want = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "sm__inst_executed.sum", "sm__inst_executed.avg.per_cycle_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.sum",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "sm__cycles_active.avg", "sm__cycles_elapsed.avg"]
print("== headline metrics")
for i, h in enumerate(hdr):
    if h in want:
        print(f"  {h:70s} {vals[i]} {units[i]}")
if "--json" in sys.argv:   # per-launch DRAM traffic for bench.py's roofline.traffic
    import json
    v = {h: float(vals[i]) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}[units[i]] for i, h in enumerate(hdr)
         if h in ("dram__bytes_read.sum", "dram__bytes_write.sum")}
    out = {"dram_bytes_per_launch": v["dram__bytes_read.sum"] + v["dram__bytes_write.sum"], "dram_bytes_read": v["dram__bytes_read.sum"],
           "dram_bytes_write": v["dram__bytes_write.sum"], "source": sys.argv[sys.argv.index("--json") + 1],
           "workload": sys.argv[sys.argv.index("--json") + 2]}
    json.dump(out, open("profiles/latest_ncu.json", "w"), indent=1)
print("== warp stall samples (pc sampling)")
st = {h: int(vals[i]) for i, h in enumerate(hdr) if h.startswith("smsp__pcsamp_warps_issue_stalled") and "not_issued" not in h and vals[i].isdigit()}
tot = sum(st.values())
for h, v in sorted(st.items(), key=lambda kv: -kv[1]):
    if v * 200 > tot:
        print(f"  {h.replace('smsp__pcsamp_warps_issue_stalled_', ''):24s} {v / tot * 100:5.1f} %")

cs = list(csv.reader(io.StringIO(ncu("--page", "source", "--print-source", "cuda,sass", "--csv"))))
addr2line, text = {}, {}
cur_file = cur = None
for r in cs:
    if len(r) >= 2 and r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
    elif len(r) >= 3 and r[0].isdigit():
        cur = (cur_file, int(r[0]))
        text[cur] = r[1].strip()[:80]
    elif len(r) >= 3 and r[0] == "" and r[2].startswith("0x"):
        addr2line[r[2]] = cur
sa = list(csv.reader(io.StringIO(ncu("--page", "source", "--csv"))))
h2 = sa[1]
ix = {h: i for i, h in enumerate(h2)}
stalls = [h for h in h2 if h.startswith("stall_") and "Not Issued" not in h]
src = open("direct_b200/csrc/ipddp_solver.h").read().split("\n")
funcs = [(n + 1, re.search(r"(\w+)\(", l).group(1)) for n, l in enumerate(src)
         if re.match(r"^(template <.*> )?(DDP_DEVICE|DDP_DEVICE_NOINLINE|DDP_HD)", l) and "(" in l]
starts = [f[0] for f in funcs]
per_f = collections.defaultdict(collections.Counter)
per_l = collections.defaultdict(collections.Counter)
ninst, nexe = collections.Counter(), collections.Counter()
for r in sa[2:]:
    ln = addr2line.get(r[0])
    if ln is None:
        key = "?"
    elif ln[0] == "ipddp_solver.h":
        k = bisect.bisect_right(starts, ln[1]) - 1
        key = funcs[k][1] if k >= 0 else "pre"
    else:
        key = ln[0]
    for s in stalls:
        try:
            per_f[key][s] += int(r[ix[s]]); per_l[ln][s] += int(r[ix[s]])
        except ValueError:
            pass
    ninst[key] += 1
    try:
        nexe[key] += int(r[ix["Instructions Executed"]])
    except ValueError:
        pass
tot = sum(sum(c.values()) for c in per_f.values())
cols = ["stall_long_sb", "stall_no_inst", "stall_wait", "stall_selected", "stall_short_sb", "stall_lg", "stall_branch_resolving", "stall_math"]
print("== stall samples per function (% of all samples)   [SASS instrs, G warp-instr executed]")
print(f"  {'function':22s} {'all':>6s} {'sass':>6s} {'Ginst':>6s} | " + " ".join(f"{c[6:][:8]:>8s}" for c in cols))
for k, c in sorted(per_f.items(), key=lambda kv: -sum(kv[1].values())):
    t = sum(c.values())
    if t * 300 < tot:
        continue
    print(f"  {k:22s} {t / tot * 100:6.2f} {ninst[k]:6d} {nexe[k] / 1e9:6.2f} | " + " ".join(f"{c[s] / tot * 100:8.2f}" for s in cols))
print(f"  total SASS instructions {sum(ninst.values())}, executed {sum(nexe.values()) / 1e9:.2f} G warp-instructions")
print("== hottest source lines")
for ln, c in sorted(per_l.items(), key=lambda kv: -sum(kv[1].values()))[:nlines]:
    t = sum(c.values())
    top = max(c.items(), key=lambda kv: kv[1])[0][6:]
    print(f"  {t / tot * 100:5.2f} %  {ln[0] if ln else '?'}:{ln[1] if ln else 0:<5d} [{top:10s}] {text.get(ln, '')}")
