"""Summarise an .ncu-rep: key raw metrics + the hottest source lines by stall samples.
usage: python scripts/ncu_summary.py gpurun_out/prof.ncu-rep [out.txt]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum ", "dram__bytes_write.sum ", "dram__bytes_read.sum,",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_tensor", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__cycles_active.avg.pct", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "smsp__cycles_active.avg ", "sm__cycles_elapsed.avg ", "smsp__inst_executed.sum ",
        "lts__t_bytes.sum ", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__average_warp", "smsp__warp_issue_stalled", "sm__mem_tensor"]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
out = []
for r in rows[2:]:
    name = dict(zip(hdr, r)).get("Kernel Name", "?")
    out.append("== kernel: %s" % name[:110])
    for h, u, v in zip(hdr, units, r):
        if any(h.startswith(k.strip()) or k.strip() in h for k in KEYS):
            if "warp_issue_stalled" in h and not h.endswith("_per_warp_active.pct"):
                continue
            out.append("  %-95s %-12s %s" % (h, u, v))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr_i = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
for hi in hdr_i:
    hdr = rows[hi]
    kname = rows[hi - 1][1] if hi > 0 and rows[hi - 1] and rows[hi - 1][0] == "Kernel Name" else "?"
    i_samp, i_src = hdr.index("# Samples"), hdr.index("Source")
    stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    body = []
    for r in rows[hi + 1:]:
        if not r or r[0] in ("Kernel Name", "Address"):
            break
        if len(r) > i_samp and r[i_samp].isdigit():
            body.append(r)
    tot = sum(int(r[i_samp]) for r in body) or 1
    out.append("== %s: hottest SASS lines by stall samples (total %d samples, %d instrs)" % (kname[:60], tot, len(body)))
    order = sorted(range(len(body)), key=lambda i: -int(body[i][i_samp]))
    for i in order[:30]:
        r = body[i]
        top = sorted(((int(r[c]) if r[c].isdigit() else 0, h) for c, h in stall_cols), reverse=True)[:2]
        out.append("  %6.2f%%  [%4d] %-70s %s" % (100 * int(r[i_samp]) / tot, i, r[i_src].strip()[:70],
                                              " ".join("%s=%d" % (h[6:], n) for n, h in top if n)))
txt = "\n".join(out)
print(txt)
if len(sys.argv) > 2:
    open(sys.argv[2], "w").write(txt + "\n")
