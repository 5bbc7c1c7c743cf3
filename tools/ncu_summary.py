#!/usr/bin/env python
"""Condense an .ncu-rep (ncu --set full) into the per-launch CSV kept under profiles/: duration, DRAM traffic, pipe and
issue utilisation, occupancy limiters, shared-memory conflicts, top stall reasons.
usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep [--traffic KERNEL_SUBSTRING=UNITS] > profiles/rNN_ncu_<what>.csv
--traffic also records dram read + write bytes of the first matching launch (which processed UNITS work items) in
profiles/ncu_traffic.json, where bench.py picks up `roofline.traffic`."""
import csv
import json
import os
import subprocess
import sys

KEYS = [("duration", "gpu__time_duration.sum"), ("regs", "launch__registers_per_thread"),
        ("warp_inst", "smsp__inst_executed.sum"), ("issue_active_pct", "smsp__issue_active.avg.pct_of_peak_sustained_active"),
        ("pipe_alu_pct", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active"),
        ("pipe_fmaheavy_pct", "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed"),
        ("pipe_fp64_inst_pct", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"),
        ("lsu_wavefronts_pct", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed"),
        ("dram_read", "dram__bytes_read.sum"), ("dram_write", "dram__bytes_write.sum"),
        ("dram_pct", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
        ("l2_pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed"),
        ("warps_active_pct", "sm__warps_active.avg.pct_of_peak_sustained_active"),
        ("occ_limit_regs", "launch__occupancy_limit_registers"), ("occ_limit_smem", "launch__occupancy_limit_shared_mem"),
        ("smem_bank_conflicts", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"),
        ("smem_wavefronts", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum")]


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    w = csv.writer(sys.stdout)
    w.writerow(["kernel", "block", "grid"] + ["%s [%s]" % (k, units[col[m]]) if m in col else k for k, m in KEYS] + ["top_stalls"])
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
    for a in [x.split("=") for i, x in enumerate(sys.argv) if i > 0 and sys.argv[i - 1] == "--traffic"]:
        for r in rows[2:]:
            if a[0] in r[col["Kernel Name"]]:
                b = sum(float(r[col[m]]) * scale[units[col[m]]] for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
                path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "ncu_traffic.json")
                if "--traffic-out" in sys.argv:
                    path = sys.argv[sys.argv.index("--traffic-out") + 1]
                d = json.load(open(path)) if os.path.exists(path) else {}
                d[a[0]] = {"units": int(a[1]), "dram_bytes": int(b), "duration_ms_under_ncu": float(r[col["gpu__time_duration.sum"]]) * {"us": 1e-3, "ms": 1.0, "ns": 1e-6, "s": 1e3}[units[col["gpu__time_duration.sum"]]],
                           "source": os.path.basename(rep)}
                json.dump(d, open(path, "w"), indent=1, sort_keys=True)
                break
    for r in rows[2:]:
        st = [(h.replace("smsp__pcsamp_warps_issue_stalled_", ""), float(r[i])) for i, h in enumerate(hdr)
              if h.startswith("smsp__pcsamp_warps_issue_stalled_") and not h.endswith("_not_issued") and r[i] not in ("", "0")]
        st.sort(key=lambda x: -x[1])
        tot = sum(v for _, v in st) or 1.0
        w.writerow([r[col["Kernel Name"]], r[col["Block Size"]], r[col["Grid Size"]]] + [r[col[m]] if m in col else "" for _, m in KEYS] +
                   [" ".join("%s=%.0f%%" % (k, 100 * v / tot) for k, v in st[:6])])


if __name__ == "__main__":
    main()
