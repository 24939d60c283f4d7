"""Summarise an `ncu --set full` report of the conv contraction launches:
    ncu -i X.ncu-rep --page raw --csv > raw.csv ; python tools/ncu_traffic.py raw.csv PAIRS out.json out.txt
Writes per-launch time / dram bytes / tensor-pipe and memory utilisation (txt) and the DRAM bytes per frame-pair of
the 12 3x3-conv launches (json, read by bench.py for roofline.traffic)."""
import csv
import json
import sys


def main():
    raw, pairs, out_json, out_txt = sys.argv[1], int(sys.argv[2]), sys.argv[3], sys.argv[4]
    rows = list(csv.reader(open(raw)))
    hdr, units = rows[0], rows[1]
    want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "sm__inst_executed_pipe_tensor.sum", "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
            "sm__throughput.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
            "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "launch__registers_per_thread",
            "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct"]
    idx = {k: hdr.index(k) for k in want if k in hdr}

    def num(r, k):
        if k not in idx:
            return float("nan")
        v = r[idx[k]].replace(",", "")
        u = units[idx[k]]
        try:
            f = float(v)
        except ValueError:
            return float("nan")
        scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1.0}
        return f * scale.get(u, 1.0)

    lines, conv_bytes, conv_n = [], 0.0, 0
    for i, r in enumerate(rows[2:]):
        name = r[idx["Kernel Name"]].split("(")[0]
        t = num(r, "gpu__time_duration.sum")
        rd, wr = num(r, "dram__bytes_read.sum"), num(r, "dram__bytes_write.sum")
        lines.append(f"#{i:2d} {name:28s} {t * 1e3:8.3f} ms  dram rd {rd / 1e9:7.3f} GB wr {wr / 1e9:7.3f} GB "
                     f"({(rd + wr) / t / 1e12:5.2f} TB/s)  sm% {num(r, 'sm__throughput.avg.pct_of_peak_sustained_elapsed'):5.1f} "
                     f"dram% {num(r, 'dram__throughput.avg.pct_of_peak_sustained_elapsed'):5.1f} "
                     f"l2% {num(r, 'lts__throughput.avg.pct_of_peak_sustained_elapsed'):5.1f} "
                     f"regs {num(r, 'launch__registers_per_thread'):.0f}")
        if 1 <= i <= 12:                      # launch 0 = first layer (K=32 contraction), 1..12 = the 3x3 convs
            conv_bytes += rd + wr
            conv_n += 1
    open(out_txt, "w").write("\n".join(lines) + "\n")
    json.dump({"dram_bytes_per_pair": conv_bytes / pairs, "pairs_profiled": pairs, "launches": conv_n,
               "source": "ncu --set full --clock-control none, tools/stage_times.py, tma::gemm_tma_px_kernel (VGG layer 1) + "
                         "tma::gemm_tma_kernel (layers 2..12), first iteration"}, open(out_json, "w"))
    print("\n".join(lines))
    print("dram bytes per pair (12 conv launches):", conv_bytes / pairs)


if __name__ == "__main__":
    main()
