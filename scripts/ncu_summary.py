"""Condense an .ncu-rep into the text summary kept under profiles/ (run here, no GPU needed):

    python scripts/ncu_summary.py gpurun_out/prof.ncu-rep [kernel-index] > profiles/rNN_name.txt
"""
import csv, io, re, subprocess, sys

rep = sys.argv[1]
which = int(sys.argv[2]) if len(sys.argv) > 2 else 0
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2 + which]
d = dict(zip(hdr, data))
u = dict(zip(hdr, units))
print(f"# {rep}: kernel {d.get('Kernel Name', '?')}  grid {d.get('Grid Size', '?')} block {d.get('Block Size', '?')}")
keys = [
    r"^gpu__time_duration\.sum$", r"^launch__registers_per_thread$", r"^launch__occupancy_limit", r"^launch__waves_per_multiprocessor$",
    r"^sm__warps_active\.avg\.pct_of_peak_sustained_active$", r"^sm__maximum_warps_per_active_cycle_pct$",
    r"^smsp__issue_active\.avg\.pct_of_peak_sustained_active$", r"^smsp__inst_executed\.sum$", r"^smsp__thread_inst_executed_per_inst_executed\.ratio$",
    r"^smsp__warps_eligible\.avg\.per_cycle_active$", r"^smsp__warps_active\.avg\.per_cycle_active$",
    r"^sm__inst_executed_pipe_(alu|fma|fmaheavy|fmalite|xu|lsu|cbu|adu|uniform|tensor\w*)\.avg\.pct_of_peak_sustained_active$",
    r"^sm__pipe_(alu|fma|fmaheavy|fmalite|xu|shared|tensor\w*)_cycles_active\.avg\.pct_of_peak_sustained_active$",
    r"^smsp__average_warps_issue_stalled_\w+_per_issue_active\.ratio$",
    r"^dram__bytes_(read|write)\.sum$", r"^l1tex__data_bank_conflicts_pipe_lsu_mem_shared\.sum$", r"^l1tex__data_pipe_lsu_wavefronts_mem_shared\.sum$",
    r"^smsp__sass_average_branch_targets_threads_uniform\.pct$", r"^sm__sass_inst_executed_op_(shared|local|global)_(ld|st)\.sum$",
    r"^smsp__inst_executed_op_branch\.sum$", r"^sm__throughput\.avg\.pct_of_peak_sustained_elapsed$",
]
for k in hdr:
    if any(re.search(p, k) for p in keys):
        v = d[k]
        if v in ("", "0", "0.000000") and "stalled" in k:
            continue
        print(f"{k} = {v} {u.get(k, '')}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
try:
    srows = list(csv.reader(io.StringIO(src)))
    hi = next(i for i, r in enumerate(srows) if r and r[0] == "Address")      # (row 0 names the kernel)
    sh = srows[hi]
    col = {n: i for i, n in enumerate(sh)}
    body = [r for r in srows[hi + 1:] if len(r) >= len(sh) - 1 and r[0].startswith("0x")]
    samp, inst = col["# Samples"], col["Instructions Executed"]
    tot = sum(float(r[samp] or 0) for r in body) or 1.0
    tot_inst = sum(float(r[inst] or 0) for r in body) or 1.0
    print(f"\n# SASS instructions: {len(body)}; warp instructions executed {tot_inst:.4g}; warp-state samples {tot:.0f}")
    print("# 30 most-sampled instructions: share of samples, share of executed instructions, avg threads, top stall, SASS")
    stalls = [c for c in sh if c.startswith("stall_") and "Not Issued" not in c]
    for r in sorted(body, key=lambda r: -float(r[samp] or 0))[:30]:
        top = max(stalls, key=lambda c: float(r[col[c]] or 0))
        print(f"{float(r[samp] or 0) / tot * 100:5.1f}%  {float(r[inst] or 0) / tot_inst * 100:5.2f}%  {r[col['Avg. Threads Executed']]:>4}  "
              f"{top[6:]:<18} {r[col['Source']].strip()[:90]}")
except Exception as e:      # the source page is optional
    print("# (source page not parsed:", e, ")")
