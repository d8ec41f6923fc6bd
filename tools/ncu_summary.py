"""Condense an `ncu --set full` report into the small per-launch table committed under profiles/:
    python tools/ncu_summary.py gpurun_out/r01_conv_v6.ncu-rep profiles/r01_conv_ncu_full_v6_summary.csv
(reads the report with `ncu -i <rep> --page raw --csv`; one column per captured launch)."""
import csv, io, subprocess, sys

METRICS = [
    "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "gpu__time_duration.sum",
    "sm__cycles_elapsed.avg", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg",
    "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum", "l1tex__data_pipe_lsu_wavefronts.avg",
    "l1tex__data_pipe_lsu_wavefronts_mem_lgds.avg", "l1tex__data_pipe_lsu_wavefronts_mem_shared.avg",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem",
]


def main(rep, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], check=True, capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    for h, i in list(col.items()):           # section-prefixed names ("TPC.TriageCompute.sm__pipe_...") -> bare metric
        if "." in h and h.split(".", 2)[-1] not in col and h[0].isupper():
            col[h.split(".", 2)[-1]] = i
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["metric", "unit"] + ["launch %d" % i for i in range(len(data))])
        w.writerow(["Kernel Name", ""] + [r[col["Kernel Name"]] for r in data])
        for m in METRICS:
            if m in col:
                w.writerow([m, units[col[m]]] + [r[col[m]] for r in data])
    print("wrote", out, "(%d launches)" % len(data))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
