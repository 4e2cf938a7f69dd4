"""Condense an .ncu-rep into the CSV of metrics we track (run here, no GPU needed):
    python scripts/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/rNN_name.csv
"""
import csv
import subprocess
import sys

WANT = ["Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed"]
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
idx = [hdr.index(w) for w in WANT if w in hdr]
out = csv.writer(sys.stdout)
out.writerow([hdr[i] for i in idx])
out.writerow([units[i] for i in idx])
for r in rows[2:]:
    out.writerow([r[i] for i in idx])

# also refresh profiles/ncu_traffic.json (per-kernel DRAM bytes per launch, largest launch of each kernel)
if len(sys.argv) > 2 and sys.argv[2] == "--traffic":
    import json, os
    names = {"mb_warp_kernel": "mb_warp", "mb_pyrdown_kernel": "mb_pyrdown", "mb_select_kernel": "mb_select", "pack_kernel": "pack",
             "weighted_group_kernel": "weighted_fuse", "mb_pyrtail_kernel": "mb_pyrtail",
             "mbw_warp_kernel": "mbw_warp", "mbs_decide_kernel": "mbs_decide", "mbx_propagate_kernel": "mbs_propagate",
             "mbs_warp_kernel": "mbs_warp", "mbs_lap_kernel": "mbs_lap", "mbc_bounds_kernel": "mbc_bounds",
             "void mbx_pyrdown_kernel<0>": "mbw_pyramid", "void mbx_pyrdown0_tma_kernel<0>": "mbw_pyramid",
             "void mbx_pyrdown_kernel<1>": "mbs_pyramid", "void mbx_pyrdown0_tma_kernel<1>": "mbs_pyramid"}
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "ncu_traffic.json")
    t = json.load(open(path)) if os.path.exists(path) else {}
    ik, ir, iw, it = hdr.index("Kernel Name"), hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum"), hdr.index("gpu__time_duration.sum")
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    for r in rows[2:]:
        k = names.get(r[ik].split("(")[0])
        if not k:
            continue
        tot = float(r[ir]) * scale[units[ir]] + float(r[iw]) * scale[units[iw]]
        if k not in t or tot > t[k]["dram_bytes_per_launch"]:
            t[k] = {"dram_bytes_per_launch": tot, "duration_us": float(r[it]), "source": os.path.basename(sys.argv[1])}
    json.dump(t, open(path, "w"), indent=1, sort_keys=True)
