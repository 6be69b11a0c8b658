"""Developer tool: `ncu --set full` report -> the small metric,unit,value CSV kept under profiles/.
   python tools/ncu_summary.py gpurun_out/X.ncu-rep profiles/rNN_ncu_<kernel>.csv [extra-metric-substring ...]"""
import csv
import subprocess
import sys

KEEP = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.sum.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.sum.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.sum.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.sum.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "launch__waves_per_multiprocessor", "sm__cycles_elapsed.avg.per_second", "lts__t_sector_hit_rate.pct"]
rep, out, extra = sys.argv[1], sys.argv[2], sys.argv[3:]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(raw.splitlines()))
names, units = rows[0], rows[1]
with open(out, "w", newline="") as f:
    w = csv.writer(f)
    for k, data in enumerate(rows[2:]):
        w.writerow(["metric", "unit", "value"] if k == 0 else ["# launch", k, ""])
        for want in KEEP + extra:
            for i, n in enumerate(names):
                if n == want or (want in extra and want in n):
                    w.writerow([n, units[i], data[i]])
print(open(out).read())
