"""profiles/r2_traffic.json from one `ncu --set full` capture of the fused kernel (raw page + source page CSV):
DRAM bytes and FP64-pipe instructions per cell and launch.
usage: python profiles/make_traffic.py raw.csv src.csv ncells_per_launch out.json "<what was captured>" """
import csv, json, re, sys
raw, src, ncells, out, what = sys.argv[1], sys.argv[2], float(sys.argv[3]), sys.argv[4], sys.argv[5]
rows = list(csv.reader(open(raw)))
hdr, units, d = rows[0], rows[1], rows[2]
def val(name):
    i = hdr.index(name)
    v = float(d[i].replace(",", ""))
    u = units[i].lower()
    return v * {"gbyte": 1e9, "mbyte": 1e6, "kbyte": 1e3, "byte": 1.0}.get(u, 1.0)
dram = val("dram__bytes_read.sum") + val("dram__bytes_write.sum")
rows = list(csv.reader(open(src)))
h = None; R = []; nsec = 0
for r in rows:
    if r and r[0] == "Address":
        h = r; nsec += 1; continue
    if nsec == 1 and h and r and r[0].startswith("0x"):
        R.append(r)
ie, isrc = h.index("Instructions Executed"), h.index("Source")
fp64 = tot = 0
for r in R:
    n = int(r[ie]); tot += n
    m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[isrc])
    op = m.group(2).split(".")[0] if m else "?"
    if op in ("DFMA", "DMUL", "DADD", "DSETP"):
        fp64 += n
json.dump({"captured": what, "cells_per_launch": ncells,
           "dram_bytes_per_cell_per_launch": dram / ncells,
           "warp_instructions_per_cell_per_launch": tot / (ncells / 32),
           "fp64_pipe_inst_per_cell_per_launch": fp64 / (ncells / 32),
           "fp64_pipe_peak_source": "profiles/micro/fp64_peak.cu on B200: 34.2 TFLOP/s = 17.1e12 DFMA/s (58.8 lanes/clk/SM at 1965 MHz)",
           "fp64_pipe_peak_tinst_s": 17.1,
           "kernel": d[hdr.index("Kernel Name")]}, open(out, "w"), indent=1)
print(open(out).read())
