"""(CPU) Per-launch summary of an `ncu --set full` report:  ncu -i X.ncu-rep --page raw --csv > raw.csv ;
    python tools/ncu_kernel_summary.py raw.csv [units_per_launch unit_name] > profiles/rNN_<what>_ncu_summary.txt
Prints duration, DRAM bytes read/written (and per unit, e.g. per frame-pair), tensor-pipe activity, issue / L1 / L2
utilisation and registers for every captured launch."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
per = float(sys.argv[2]) if len(sys.argv) > 2 else None
unit = sys.argv[3] if len(sys.argv) > 3 else "unit"
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
W = [("gpu__time_duration.sum", "duration"), ("dram__bytes_read.sum", "dram read"), ("dram__bytes_write.sum", "dram write"),
     ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active %"),
     ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots active %"),
     ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "L1/TEX throughput %"),
     ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput %"),
     ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput %"),
     ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
     ("launch__registers_per_thread", "registers/thread"), ("launch__grid_size", "grid"), ("launch__block_size", "block")]


def to_bytes(v, u):
    v = float(v.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(u, 1)


for r in rows[2:]:
    print(r[idx["Kernel Name"]])
    tot = 0.0
    for key, name in W:
        if key not in idx:
            continue
        v, u = r[idx[key]], units[idx[key]]
        line = f"    {name:26s} {v:>16s} {u}"
        if key.startswith("dram__bytes"):
            b = to_bytes(v, u)
            tot += b
            if per:
                line += f"   = {b / per / 1e6:9.1f} MB per {unit}"
        print(line)
    if per:
        print(f"    {'dram read + write':26s} {tot / 1e6:16.1f} Mbyte   = {tot / per / 1e6:9.1f} MB per {unit}")
