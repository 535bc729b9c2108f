#!/usr/bin/env python3
"""Summarise an .ncu-rep (one kernel): headline metrics, stall reasons, opcode mix.  python tools/ncu_summary.py rep [weights]"""
import collections, csv, io, re, subprocess, sys
rep = sys.argv[1]
weights = float(sys.argv[2]) if len(sys.argv) > 2 else None
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, vals = rows[0], rows[2]
m = dict(zip(hdr, vals))
def g(k):
    try: return float(m[k].replace(",", ""))
    except Exception: return float("nan")
print("kernel:", m.get("Kernel Name", "")[:90], "grid", m.get("Grid Size"), "block", m.get("Block Size"))
print(f"duration {g('gpu__time_duration.sum'):.2f} us | dram read {g('dram__bytes_read.sum'):.2f} {m.get('dram__bytes_read.sum','')} write {g('dram__bytes_write.sum'):.2f} | "
      f"dram% {g('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'):.1f} | regs {g('launch__registers_per_thread'):.0f} | dyn smem {g('launch__shared_mem_per_block_dynamic'):.1f}")
print(f"issue active {g('smsp__issue_active.avg.pct_of_peak_sustained_active'):.1f}% | alu {g('sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active'):.1f}% "
      f"fma {g('sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active'):.1f}% | warps active {g('sm__warps_active.avg.pct_of_peak_sustained_active'):.1f}% | "
      f"smem bank conflicts {g('l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum'):.0f} / wavefronts {g('l1tex__data_pipe_lsu_wavefronts_mem_shared.sum'):.0f}")
st = {h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""): g(h) for h in hdr
      if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")}
print("stall cycles per issued instr:", ", ".join(f"{k} {v:.2f}" for k, v in sorted(st.items(), key=lambda kv: -kv[1]) if v > 0.03), "| total", round(sum(st.values()), 2))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
h = rows[hi]
S, E = h.index("Source"), h.index("Instructions Executed")
ops, tot = collections.Counter(), 0
for r in rows[hi + 1:]:
    try: n = int(r[E])
    except Exception: continue
    t = r[S].split()
    op = t[1] if t and t[0].startswith("@") else (t[0] if t else "?")
    ops[re.sub(r"\..*", "", op)] += n
    tot += n
print("warp instructions", tot, (f"= {tot * 32 / weights:.3f} thread-instr/weight" if weights else ""))
print("  ".join(f"{o} {n * 32 / weights:.3f}" if weights else f"{o} {n}" for o, n in ops.most_common(16)))
