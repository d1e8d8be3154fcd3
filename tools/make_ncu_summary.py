#!/usr/bin/env python
"""gpurun_out/r2_hot.ncu-rep (+ the U-Net step launch list) -> profiles/r2_ncu_hot_kernels.csv and
profiles/r2_ncu_summary.json (the ncu-only numbers bench.py quotes: DRAM bytes of one U-Net step, tensor-pipe
activity of the attention kernel).  Run HERE (no GPU needed): python tools/make_ncu_summary.py"""
import csv
import json
import os
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep = sys.argv[1] if len(sys.argv) > 1 else os.path.join(REPO, "gpurun_out", "r2_hot.ncu-rep")
out = {}
if os.path.exists(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    want = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
            "sm__throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread",
            "sm__warps_active.avg.pct_of_peak_sustained_active"]
    want += [h for h in hdr if "pipe_tensor" in h and h not in want]
    idx = [hdr.index(k) for k in want if k in hdr]
    names = ["attn L2560 self", "attn L640 self", "attn L160 self", "attn cross 77+16 @L2560", "conv3x3 320 @40x64",
             "conv3x3 640 @20x32", "conv3x3 1280 @5x8", "split-K reduce", "linear 40960x320->320",
             "GEGLU 320->2560", "GroupNorm+SiLU 16x2560x320", "align loop (3 iterations)"]
    path = os.path.join(REPO, "profiles", "r2_ncu_hot_kernels.csv")
    with open(path, "w") as f:
        f.write("# ncu --set full --clock-control none, one launch per hot kernel (tools/prof_hot.py); cold cache, serialised\n")
        f.write("launch," + ",".join(f"{hdr[i]}[{units[i]}]" if units[i] else hdr[i] for i in idx) + "\n")
        labels, li = [], 0
        kcol = hdr.index("Kernel Name")
        for r in data:   # the split-K reduce row only exists when the library split the 5x8 conv (not with its cost model's BN=64)
            if li < len(names) and names[li] == "split-K reduce" and "splitk" not in r[kcol]:
                li += 1
            labels.append(names[li] if li < len(names) else str(li))
            li += 1
        for n, r in enumerate(data):
            f.write(labels[n] + "," + ",".join('"' + r[i][:60] + '"' if hdr[i] == "Kernel Name"
                                                                            else r[i].replace(",", "") for i in idx) + "\n")
    print("wrote", path)
    tcol = [h for h in hdr if h.startswith("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active")] or \
           [h for h in hdr if "pipe_tensor" in h and "pct_of_peak_sustained_active" in h]
    if tcol and data:
        j = hdr.index(tcol[0])
        for key, n in (("attn_L2560_pipe_tensor_pct", 0), ("attn_L640_pipe_tensor_pct", 1), ("attn_L160_pipe_tensor_pct", 2),
                       ("attn_cross_pipe_tensor_pct", 3), ("conv3x3_320_pipe_tensor_pct", 4), ("conv3x3_640_pipe_tensor_pct", 5)):
            if n < len(data):
                try:
                    out[key] = float(data[n][j].replace(",", ""))
                except ValueError:
                    pass
        out["pipe_tensor_metric"] = tcol[0]
tj = os.path.join(REPO, "profiles", "r2_unet_step_traffic.json")
if os.path.exists(tj):
    t = json.load(open(tj))
    out["unet_step_dram_bytes"] = t.get("dram_bytes")
    out["unet_step_sum_kernel_ms"] = t.get("sum_kernel_ms")
    out["unet_step_launches"] = t.get("launches")
json.dump(out, open(os.path.join(REPO, "profiles", "r2_ncu_summary.json"), "w"), indent=1, sort_keys=True)
print(json.dumps(out, indent=1))
