#!/usr/bin/env python
"""`ncu -i report --page raw --csv` -> compact per-launch table of the metrics the roofline discussion uses."""
import csv, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, data = rows[0], rows[1], rows[2:]
want = [("Kernel Name", "kernel"), ("Grid Size", "grid"), ("Block Size", "block"), ("gpu__time_duration.sum", "duration"),
        ("sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active", "tensor_pipe_pct"),
        ("sm__inst_executed_pipe_tensor.sum", "tensor_inst"),
        ("dram__bytes_read.sum", "dram_read"), ("dram__bytes_write.sum", "dram_write"),
        ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"),
        ("lts__t_bytes.sum", "l2_bytes"), ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2_pct"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_pct"),
        ("launch__registers_per_thread", "regs"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active_pct")]
idx = [(hdr.index(k), n) for k, n in want if k in hdr]
print(",".join(f"{n}[{units[i]}]" if units[i] else n for i, n in idx))
for r in data:
    print(",".join('"' + r[i][:70] + '"' if n == "kernel" else r[i].replace(",", "") for i, n in idx))
