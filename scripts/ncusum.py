"""Print the roofline-relevant counters of an .ncu-rep (first profiled launch).

    python scripts/ncusum.py REPORT.ncu-rep
    python scripts/ncusum.py REPORT.ncu-rep --traffic k_create_rays --units 33177600 [--out profiles/kernel_traffic.json]

--traffic records dram__bytes_read.sum + dram__bytes_write.sum of the captured launch per unit (ray / splat) together with
the hash of the kernel's sources (bench.kernel_source_hash) in a JSON file that bench.py reads for `roofline.traffic`;
a capture taken on older sources is detected there as stale.
"""
import csv
import json
import os
import subprocess
import sys

rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
r = list(csv.reader(out.splitlines()))
hdr, units, vals = r[0], r[1], r[2]
want = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__inst_executed_pipe_fma.sum", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.sum", "sm__inst_executed_pipe_xu.sum", "sm__inst_executed_pipe_fp64.sum", "sm__inst_executed_pipe_lsu.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sectors_op_red.sum", "lts__t_sectors_op_atom.sum",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__sass_thread_inst_executed_op_ffma_pred_on.sum",
        "smsp__sass_thread_inst_executed_op_fmul_pred_on.sum", "smsp__sass_thread_inst_executed_op_fadd_pred_on.sum", "sm__cycles_elapsed.max",
        "smsp__warps_eligible.avg.per_cycle_active", "sm__sass_thread_inst_executed_op_ffma_pred_on.sum", "sm__sass_thread_inst_executed_op_fmul_pred_on.sum",
        "sm__sass_thread_inst_executed_op_fadd_pred_on.sum",
        # the splat accumulate: reductions as the L2 slices see them (sectors, share of the slices' peak, hit / miss), and on the way there
        "lts__t_sectors_srcunit_tex_op_red.sum", "lts__t_sectors_srcunit_tex_op_red.sum.pct_of_peak_sustained_elapsed",
        "lts__t_sectors_srcunit_tex_op_red.sum.per_second", "lts__t_sectors_srcunit_tex_op_red_lookup_hit.sum",
        "lts__t_sectors_srcunit_tex_op_red_lookup_miss.sum", "lts__t_sectors_srcunit_tex_op_atom.sum", "lts__t_requests_srcunit_tex_op_red.sum",
        "l1tex__m_l1tex2xbar_write_sectors_mem_global_op_red.sum", "l1tex__m_l1tex2xbar_write_sectors_mem_global_op_red.sum.pct_of_peak_sustained_elapsed",
        "l1tex__t_requests_pipe_lsu_mem_global_op_red.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_atom.sum",
        "l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "lts__t_sectors.sum.pct_of_peak_sustained_elapsed"]
for i, h in enumerate(hdr):
    if h in want or (h.startswith("smsp__average_warps_issue_stalled") and h.endswith("_per_issue_active.ratio")) or \
            (h.startswith("smsp__average_warp_latency_issue_stalled") and h.endswith(".ratio")):
        print(f"{h} = {vals[i]} {units[i]}")


def _num(name):
    i = hdr.index(name)
    v = float(vals[i].replace(",", ""))
    u = units[i].lower()
    return v * {"byte": 1.0, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "tbyte": 1e12}.get(u, 1.0)


if "--traffic" in sys.argv:
    key = sys.argv[sys.argv.index("--traffic") + 1]
    n_units = float(sys.argv[sys.argv.index("--units") + 1])
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = sys.argv[sys.argv.index("--out") + 1] if "--out" in sys.argv else os.path.join(root, "profiles", "kernel_traffic.json")
    sys.path.insert(0, root)
    import bench

    data = {}
    if os.path.exists(out):
        with open(out) as f:
            data = json.load(f)
    rd, wr = _num("dram__bytes_read.sum"), _num("dram__bytes_write.sum")
    data[key] = {"dram_bytes_per_ray" if "ray" in key else "dram_bytes_per_unit": (rd + wr) / n_units, "dram_bytes_read": rd, "dram_bytes_write": wr,
                 "units": n_units, "kernel": vals[hdr.index("Kernel Name")], "source": "ncu --set full capture " + os.path.basename(rep),
                 "kernel_source_hash": bench.kernel_source_hash()}
    with open(out, "w") as f:
        json.dump(data, f, indent=1)
    print("wrote", out, data[key])
