#!/usr/bin/env python
"""profiles/r02_multi_gpu/*.json (bench.py --gpus N lines) -> profiles/r02_multi_gpu.md"""
import glob
import json
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rows = []
for f in sorted(glob.glob(os.path.join(ROOT, "profiles", "r02_multi_gpu", "*.json"))):
    d = json.loads(open(f).read().strip().splitlines()[-1])
    m = re.search(r"beta=([0-9.]+)", d["config"]["workload"])
    rows.append({
        "file": os.path.basename(f), "n": d["n_gpus"], "scaling": d["scaling"], "beta": float(m.group(1)) if m else None,
        "ms": d["ms_per_step"], "gedges": d["value"] / 1e9, "frac": d["roofline"]["frac"], "tbps": d["roofline"]["achieved"] / 1e3,
        "parity": d["parity"], "e2e_ms": d["e2e"]["ms_per_step"], "pcie": d["e2e"]["pcie_GBps_per_rank_each_way"],
        "anchor": d.get("anchor_local_only", {}), "workload": d["config"]["workload"],
    })
rows.sort(key=lambda r: (r["scaling"], r["beta"], r["n"]))
out = ["# Multi-GPU session, round 2 (one box, NVLink 5 / NVSwitch; `tools/run_multi_gpu.sh`, `bench.py --gpus N`)", "",
       "Per-rank work: weak = 1/8 of a papers100M-shaped graph per GPU (13 882 494 rows, 201 960 734 edges, hidden = 128, unweighted);",
       "strong = the whole papers100M-shaped graph (111 M rows, 1.6 B edges) split N ways.  beta = probability that a column is drawn",
       "over the whole graph instead of the rank's own node range (remote-edge fraction = beta (N-1)/N).  Remote feature rows are",
       "read from the owner's HBM over NVLink INSIDE the SpMM kernel (`cogdl_b200_spmm_csr_f32_peers`); no collective on the data path.",
       "`anchor` = the same shards with every column local, timed in the same job on all ranks at once (the like-for-like 1-GPU rate);",
       "`eff` = anchor ms / measured ms.  `parity` = rows checked against the oracle inside the run (all ranks), worst element error.", "",
       "| scaling | beta | GPUs | ms/step | G edges/s (all ranks) | per-rank TB/s (algorithmic) | frac of HBM peak | anchor ms | eff vs anchor | parity rows / max rel / ok | e2e ms | PCIe GB/s per rank each way |",
       "|---|---|---|---|---|---|---|---|---|---|---|---|"]
for r in rows:
    a = r["anchor"]
    out.append(f"| {r['scaling']} | {r['beta']} | {r['n']} | {r['ms']:.2f} | {r['gedges']:.1f} | {r['tbps']:.2f} | {r['frac']:.3f} | "
               f"{a.get('ms_per_step', float('nan')):.2f} | {a.get('efficiency_vs_anchor', float('nan')):.3f} | "
               f"{r['parity']['rows']} / {r['parity']['max_rel']:.1e} / {r['parity']['ok']} | {r['e2e_ms']:.1f} | {r['pcie']:.1f} |")
out += ["", "1-GPU anchor of the weak curve measured by the N = 1 bench (`others.C5_shard_1gpu`, `profiles/r02m_bench_n1.json`): "
        "16.68 ms/step = 12.1 G edges/s.", ""]
open(os.path.join(ROOT, "profiles", "r02_multi_gpu.md"), "w").write("\n".join(out))
print("\n".join(out))
