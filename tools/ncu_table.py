#!/usr/bin/env python3
"""One line per profiled launch from an `ncu --page raw --csv` export (or an .ncu-rep): duration, DRAM bytes and %, issue activity,
top stall reasons.  python tools/ncu_table.py <raw.csv | rep.ncu-rep> [peak_GBps]"""
import csv, io, subprocess, sys
src = sys.argv[1]
peak = float(sys.argv[2]) if len(sys.argv) > 2 else 6575.4
text = open(src).read() if src.endswith(".csv") else subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(text)))
hdr = rows[0]
def unit_scale(u):
    return {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1e3, "us": 1.0, "ns": 1e-3, "s": 1e6}.get(u, 1.0)
units = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
def val(r, k, scale=True):
    try:
        v = float(r[ix[k]].replace(",", ""))
    except Exception:
        return float("nan")
    return v * unit_scale(units[ix[k]]) if scale else v
stall_keys = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
print(f"{'kernel':58s} {'grid':>12s} {'blk':>4s} {'us':>7s} {'rdMB':>8s} {'wrMB':>6s} {'GB/s':>7s} {'ofpk':>5s} {'dram%':>6s} {'issue%':>6s} {'regs':>4s} {'smemKB':>6s}  top stalls (cycles per issued instr)")
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    name = r[ix["Kernel Name"]].replace("nt::b200::<unnamed>::", "").replace("(nt::b200::<unnamed>::KqParams)", "")[:58]
    us = val(r, "gpu__time_duration.sum")
    rd, wr = val(r, "dram__bytes_read.sum"), val(r, "dram__bytes_write.sum")
    st = sorted(((val(r, k, False), k.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")) for k in stall_keys), reverse=True)[:4]
    gbs = (rd + wr) / (us * 1e-6) / 1e9 if us == us and us > 0 else float("nan")
    print(f"{name:58s} {r[ix['Grid Size']]:>12s} {r[ix['Block Size']].split(',')[0].strip('( '):>4s} {us:7.2f} {rd / 1e6:8.2f} {wr / 1e6:6.2f} {gbs:7.0f} {gbs / peak:5.2f} "
          f"{val(r, 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', False):6.1f} {val(r, 'smsp__issue_active.avg.pct_of_peak_sustained_active', False):6.1f} "
          f"{val(r, 'launch__registers_per_thread', False):4.0f} {val(r, 'launch__shared_mem_per_block_dynamic') / 1024:6.1f}  " + ", ".join(f"{n} {v:.2f}" for v, n in st))
