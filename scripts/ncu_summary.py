"""Per-launch summary of an `ncu --set full` report: duration, DRAM bytes, SM / issue activity.
usage: ncu_summary.py report.ncu-rep > profiles/<tag>_ncu_full_summary.txt"""
import csv, subprocess, sys

rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units, data = rows[0], rows[1], rows[2:]
col = {n: i for i, n in enumerate(hdr)}
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "lts__t_sector_hit_rate.pct",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "launch__shared_mem_per_block_dynamic",
        # the L1 / shared-memory data pipe: LSU wavefronts (shared ld / st, shuffles, local / global, TMEM loads)
        # and the tensor core's operand reads share it -- the resource that bounds the pipelined kernels
        "l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum", "l1tex__data_pipe_tc_wavefronts_mem_shared.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"]
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-9, "us": 1e-6, "ms": 1e-3, "second": 1.0, "s": 1.0}
for r in data:
    name = r[col["Kernel Name"]]
    print(f"== {name[:70]}  grid {r[col['Grid Size']]} block {r[col['Block Size']]}")
    vals = {}
    for w in want:
        if w in col:
            u = units[col[w]]
            print(f"   {w:72s} {r[col[w]]:>16s} {u}")
            try:
                vals[w] = float(r[col[w]].replace(",", "")) * scale.get(u, 1.0)
            except ValueError:
                pass
    try:
        lsu = vals["l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed"]
        tcw = vals["l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed"]
        print(f"   => L1 data pipe: LSU {lsu:.1f} % + tensor-core operands {tcw:.1f} % = {lsu + tcw:.1f} % of its peak")
    except KeyError:
        pass
    try:
        tot = vals["dram__bytes_read.sum"] + vals["dram__bytes_write.sum"]
        print(f"   => DRAM traffic {tot/1e6:.1f} MB per launch, {tot/vals['gpu__time_duration.sum']/1e9:.0f} GB/s under ncu")
    except KeyError:
        pass
