"""`ncu --set full` report -> the small `metric,unit,value` CSV kept under profiles/ (one kernel launch per report):
  python tools/ncu_extract.py gpurun_out/x.ncu-rep profiles/r02_x.ncu.csv
Keeps duration, DRAM / L2 / L2->SM bytes, pipe utilisation (tensor, XU = MUFU, FMA), issue / occupancy, launch geometry, the
per-reason warp-stall ratios, and the ten SASS lines with the most stall samples (source page)."""
import csv
import io
import subprocess
import sys

rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
KEEP = ("Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "lts__t_bytes.sum", "l1tex__m_xbar2l1tex_read_bytes.sum", "sm__cycles_elapsed.avg", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "launch__grid_size", "launch__block_size",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__cluster_size")
with open(out, "w", newline="") as fh:
    w = csv.writer(fh)
    w.writerow(["metric", "unit", "value"])
    for name in KEEP:
        if name in hdr:
            i = hdr.index(name)
            w.writerow([name, units[i], vals[i]])
    for i, name in enumerate(hdr):
        if name.startswith("smsp__average_warps_issue_stalled_") and name.endswith("_per_issue_active.ratio"):
            w.writerow([name, units[i], vals[i]])
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
    srows = list(csv.reader(io.StringIO(src)))
    if len(srows) > 2 and "# Samples" in srows[1]:
        h = srows[1]
        ix = {n: k for k, n in enumerate(h)}
        data = [r for r in srows[2:] if len(r) == len(h)]
        tot = sum(int(r[ix["# Samples"]] or 0) for r in data) or 1
        for r in sorted(data, key=lambda r: -int(r[ix["# Samples"]] or 0))[:10]:
            st = {k: int(r[ix[k]] or 0) for k in h if k.startswith("stall_") and "Not Issued" not in k}
            top = max(st, key=st.get)
            w.writerow([f"top_stall_sass: {r[ix['Source']].strip()[:80]}", "% of samples", f"{100 * int(r[ix['# Samples']]) / tot:.1f} ({top})"])
print(open(out).read()[:600])
