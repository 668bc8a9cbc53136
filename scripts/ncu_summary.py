"""Selected raw metrics of an .ncu-rep as a small CSV (one block per profiled launch) for profiles/.
usage: python scripts/ncu_summary.py gpurun_out/x.ncu-rep profiles/x_summary.csv"""
import csv
import subprocess
import sys

KEEP = ("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__occupancy_limit", "launch__shared_mem_per_block",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active",
        "sm__inst_executed_pipe_fp64", "smsp__inst_executed.sum", "smsp__average_warps_issue_stalled",
        "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum", "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum",
        "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum", "l1tex__data_pipe_lsu_wavefronts.avg.pct",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "pipe_tensor_subpipe_dmma")
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
with open(sys.argv[2], "w") as f:
    w = csv.writer(f)
    for vals in rows[2:]:
        d = dict(zip(hdr, vals))
        w.writerow(["kernel", "", d.get("Kernel Name", "")])
        w.writerow(["metric", "unit", "value"])
        for h, u in zip(hdr, units):
            if any(k in h for k in KEEP) and "not_issued" not in h and d[h] not in ("", "n/a"):
                if "issue_stalled" in h:
                    try:
                        if float(d[h]) < 0.05:
                            continue
                    except ValueError:
                        pass
                w.writerow([h, u, d[h]])
