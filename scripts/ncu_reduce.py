"""Dev helper: reduce an .ncu-rep (one kernel launch) to the (metric, unit, value) lines kept under profiles/.
usage: python scripts/ncu_reduce.py report.ncu-rep > profiles/<name>_ncu_raw.csv"""
import csv
import io
import re
import subprocess
import sys

KEEP = re.compile(r"^(ID|Kernel Name|Block Size|Grid Size|gpu__time_duration|launch__|dram__bytes_(read|write)\.sum($|\.per_second)|"
                  r"dram__throughput|sm__issue_active|smsp__issue_active|sm__inst_executed_pipe_(fp64|tensor|lsu|alu|fma)|"
                  r"sm__pipe_fp64|sm__warps_active|smsp__warps_active|smsp__average_warps_issue_stalled|smsp__inst_executed\.sum|"
                  r"smsp__thread_inst_executed_per_inst_executed|sass__inst_executed_(local|global|shared)|"
                  r"l1tex__data_pipe_lsu_wavefronts_mem_shared|l1tex__data_bank_conflicts_pipe_lsu_mem_shared|"
                  r"l1tex__t_sector_hit_rate|lts__t_sector_hit_rate|sm__throughput|gpc__cycles_elapsed\.max|sm__cycles_active\.avg|"
                  r"smsp__pcsamp_warps_issue_stalled|sm__sass_thread_inst_executed_op_d|smsp__sass_thread_inst_executed_op_d)")
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
w = csv.writer(sys.stdout)
for name, unit, val in zip(rows[0], rows[1], rows[2]):
    if KEEP.match(name):
        w.writerow([name, unit, val])
