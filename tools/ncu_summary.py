"""Summarise an ncu report: headline metrics + hottest SASS lines (run where ncu is installed)."""
import csv, subprocess, sys, io
rep = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw))); hdr = rows[0]
want = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__average_warp_latency_per_inst_issued.ratio", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "sass__inst_executed_local_loads", "sass__inst_executed_local_stores", "smsp__cycles_active.avg", "sm__cycles_active.avg", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum"]
for r in rows[2:]:
    print("== kernel:", r[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "")
    for i, h in enumerate(hdr):
        if h in want or ("issue_stalled" in h and h.endswith("per_issue_active.ratio")):
            try:
                if abs(float(r[i].replace(",", ""))) < 0.005: continue
            except ValueError: pass
            print(f"  {h} [{rows[1][i]}] = {r[i]}")
sass = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(sass)))
h = next(i for i, r in enumerate(rows) if r and r[0] == "Address"); hdr = rows[h]; data = [r for r in rows[h + 1:] if len(r) == len(hdr)]
ix = {k: i for i, k in enumerate(hdr)}
def f(r, k):
    try: return float(r[ix[k]].replace(",", ""))
    except Exception: return 0.0
tot = sum(f(r, "# Samples") for r in data)
print(f"== hottest SASS ({int(tot)} samples, {int(sum(f(r,'Instructions Executed') for r in data))} warp instructions)")
for r in sorted(data, key=lambda r: -f(r, "# Samples"))[:topn]:
    st = {k[6:]: int(f(r, k)) for k in hdr if k.startswith("stall_") and "Not Issued" not in k and f(r, k) > 0.1 * f(r, "# Samples")}
    print(f"  {r[ix['Address']][-5:]} {r[ix['Source']][:56]:56s} {100*f(r,'# Samples')/tot:5.1f}% exec={int(f(r,'Instructions Executed'))} thr={r[ix['Avg. Threads Executed']]} {st}")
