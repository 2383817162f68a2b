"""Compact per-kernel summary of an .ncu-rep (raw page): python tools/ncu_summary.py file.ncu-rep [out.md]"""
import csv, io, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
idx = {h: i for i, h in enumerate(hdr)}
M = [("gpu__time_duration.sum", "time"), ("launch__grid_size", "grid"), ("launch__registers_per_thread", "regs"),
     ("launch__shared_mem_per_block_dynamic", "dyn smem"),
     ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor pipe % (elapsed)"),
     ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe % (active)"),
     ("dram__bytes_read.sum", "dram read"), ("dram__bytes_write.sum", "dram write"),
     ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram % of peak"),
     ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 % of peak"),
     ("lts__t_bytes.sum", "L2 bytes"),
     ("l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "smem (tensor/TMA) wavefront %"),
     ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
     ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %")]
out = []
for d in data:
    name = d[idx["Kernel Name"]]
    out.append(f"### {name[:90]}  grid={d[idx['Grid Size']]} block={d[idx['Block Size']]}")
    for key, label in M:
        if key in idx:
            out.append(f"- {label}: {d[idx[key]]} {units[idx[key]]}")
    out.append("")
text = "\n".join(out)
print(text)
if len(sys.argv) > 2:
    open(sys.argv[2], "w").write(text)
