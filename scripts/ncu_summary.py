"""Summarise an .ncu-rep (run here, no GPU needed): python scripts/ncu_summary.py <rep> [regex]"""
import csv, re, subprocess, sys
rep = sys.argv[1]
pat = re.compile(sys.argv[2] if len(sys.argv) > 2 else
                 r"^(gpu__time_duration.sum|dram__bytes_(read|write).sum$|lts__t_bytes.sum$|sm__cycles_elapsed.max$|sm__cycles_active.avg$|"
                 r"launch__(registers_per_thread|grid_size|block_size|shared_mem_per_block_dynamic)$|"
                 r"sm__warps_active.avg.pct_of_peak_sustained_active|sm__throughput.avg.pct_of_peak_sustained_elapsed|"
                 r"gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed|"
                 r".*pipe_tensor.*(pct|cycles_active).*|sm__inst_executed_pipe_(alu|fma|fmaheavy|lsu|uniform|tmem|tc).*sum$|"
                 r"sm__inst_executed_pipe_.*pct_of_peak_sustained_active$|smsp__inst_executed.sum$|"
                 r"l1tex__data_pipe_lsu_wavefronts_mem_shared.sum$|smsp__average_warp.*issue_stalled.*ratio$|"
                 r"sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_(active|elapsed)$)")
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    name = r[hdr.index("Kernel Name")]
    print("=== ", name[:110])
    for i, h in enumerate(hdr):
        if pat.match(h) and r[i] not in ("", "0"):
            print(f"  {h:95s} {r[i]:>16s} {units[i]}")
