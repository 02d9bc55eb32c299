"""Text summary of one kernel of an ncu report for profiles/: selected raw metrics, warp-state mix, per-device-function table.

    python tools/ncu_summary.py prof.ncu-rep libbsgp.so <kernel substring> "<header text>" > profiles/xxx.txt
"""
import csv, io, subprocess, sys
rep, so, kern, header = sys.argv[1:5]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, val = rows[0], rows[1], rows[2]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "sass__inst_executed_global_loads",
        "sass__inst_executed_global_stores", "sass__inst_executed_shared_loads", "sass__inst_executed_shared_stores",
        "sass__inst_executed_local_loads", "sass__inst_executed_local_stores", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__cluster_size",
        "launch__shared_mem_per_block_dynamic", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]
print(header)
print("\n--- selected raw metrics (ncu -i prof.ncu-rep --page raw --csv) ---")
for w in want:
    if w in hdr:
        i = hdr.index(w)
        print(f"{w} [{units[i]}] = {val[i]}")
st = {}
for i, h in enumerate(hdr):
    if h.startswith("smsp__pcsamp_warps_issue_stalled_") and not h.endswith("_not_issued"):
        try:
            st[h.replace("smsp__pcsamp_warps_issue_stalled_", "")] = float(val[i].replace(",", ""))
        except ValueError:
            pass
tot = sum(st.values()) or 1.0
print("warp-state samples: " + ", ".join(f"{k} {100 * v / tot:.1f}%" for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:9]))
print("\n--- per device function (tools/ncu_by_function.py) ---")
print(subprocess.run([sys.executable, __file__.replace("ncu_summary.py", "ncu_by_function.py"), rep, so, kern], capture_output=True, text=True).stdout)
