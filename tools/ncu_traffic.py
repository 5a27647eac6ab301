#!/usr/bin/env python
"""Summarise an `ncu --set full` capture (.ncu-rep) of one kernel into the small JSON that bench.py
reads for `roofline.traffic`, plus a text summary for profiles/.

  python tools/ncu_traffic.py gpurun_out/prof.ncu-rep --kernel stream_kernel \
        --launched-as "cogdl_b200::stream_kernel<float4,NV=1,weighted,SRC_ONE,U=4,MINB=5>" \
        --out profiles/r02_spmm_traffic.json [--summary profiles/r02_spmm_ncu_summary.txt]

`--launched-as` is the string cogdl_b200_last_kernel() reported for the run that was profiled (bench.py
prints it as roofline.kernel); bench.py only uses the traffic figure when it matches the kernel it
launched.  Values are per launch: the mean over the captured launches of that kernel.
"""
import argparse
import csv
import io
import json
import re
import subprocess
import sys

METRICS = ["dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum", "lts__t_sector_hit_rate.pct",
           "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
           "launch__registers_per_thread", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
           "lts__t_bytes.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "launch__occupancy_limit_registers",
           "launch__occupancy_limit_shared_mem", "sm__maximum_warps_per_active_cycle_pct", "launch__shared_mem_per_block_dynamic",
           "l1tex__t_bytes.sum", "smsp__warps_eligible.avg.per_cycle_active"]
UNIT_SCALE = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6,
              "usecond": 1, "nsecond": 1e-3, "msecond": 1e3, "second": 1e6}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("rep")
    ap.add_argument("--kernel", required=True, help="regex matched against the kernel name")
    ap.add_argument("--launched-as", default="")
    ap.add_argument("--out", required=True)
    ap.add_argument("--summary")
    a = ap.parse_args()
    raw = subprocess.run(["ncu", "-i", a.rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr_i = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    hdr, units = rows[hdr_i], rows[hdr_i + 1]
    name_col = hdr.index("Kernel Name")
    pick = [r for r in rows[hdr_i + 2:] if len(r) == len(hdr) and re.search(a.kernel, r[name_col])]
    if not pick:
        sys.exit(f"no kernel matching {a.kernel!r} in {a.rep}")
    out = {"ncu_kernel_name": pick[0][name_col], "launched_as": a.launched_as, "launches": len(pick), "report": a.rep}
    for m in METRICS:
        if m not in hdr:
            continue
        c = hdr.index(m)
        vals = []
        for r in pick:
            try:
                vals.append(float(r[c].replace(",", "")) * UNIT_SCALE.get(units[c], 1))
            except ValueError:
                pass
        if vals:
            out[m] = sum(vals) / len(vals)
    out["dram_bytes_read"] = out.get("dram__bytes_read.sum")
    out["dram_bytes_write"] = out.get("dram__bytes_write.sum")
    out["duration_us"] = out.get("gpu__time_duration.sum")
    with open(a.out, "w") as f:
        json.dump(out, f, indent=1)
    if a.summary:
        with open(a.summary, "w") as f:
            f.write(f"# {a.rep}: kernel /{a.kernel}/ ({len(pick)} launches), per-launch means\n")
            for k, v in out.items():
                f.write(f"{k:70s} {v}\n")
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
