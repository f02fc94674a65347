"""Condense an .ncu-rep (ncu --set full) into the per-kernel summary CSV committed under profiles/:
   python tools/ncu_summary.py gpurun_out/x.ncu-rep profiles/rNN_name_ncu_summary.csv"""
import csv, io, subprocess, sys
COLS = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_tc_wavefronts_mem_shared.sum",
        "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "sm__cycles_elapsed.max", "smsp__inst_executed.sum",
        "lts__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__m_xbar2l1tex_read_bytes.sum", "sm__inst_executed.avg.per_cycle_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active"]
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, body = rows[0], rows[1], rows[2:]
idx = [hdr.index(c) if c in hdr else -1 for c in COLS]
with open(sys.argv[2], "w", newline="") as f:
    wr = csv.writer(f)
    wr.writerow(COLS)
    wr.writerow([units[i] if i >= 0 else "" for i in idx])
    seen = set()
    for r in body:
        name = r[idx[0]]
        if name in seen:          # first captured launch of each kernel
            continue
        seen.add(name)
        wr.writerow([r[i] if i >= 0 else "" for i in idx])
print(open(sys.argv[2]).read())
