#!/bin/bash
# Multi-GPU measurement session (one box, N GPUs):  tools/run_multi_gpu.sh N [tag]
#   tools/run_multi_gpu.sh N tag "weak_b005 strong_b005" runs a subset;
#   weak scaling (default beta 0.05), the locality sweep beta in {0.25, 1.0}, strong scaling (papers100M / N),
#   and the gloo-free NCCL parity that bench.py runs inside every configuration.  JSON lines land in gpurun_out/.
set -u
N=${1:-2}
TAG=${2:-r02}
OUT=gpurun_out
mkdir -p $OUT
run() {  # name, extra args...
  local name=$1; shift
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --gpus $N "$@" > $OUT/${TAG}_n${N}_${name}.json 2> $OUT/${TAG}_n${N}_${name}.err
  echo "== $name rc=$? $(head -c 400 $OUT/${TAG}_n${N}_${name}.json | cut -c1-300)"
  tail -c 300 $OUT/${TAG}_n${N}_${name}.err
}
CONFIGS=${3:-"weak_b005 weak_b025 weak_b100 strong_b005"}
for c in $CONFIGS; do
  case $c in
    weak_b005) run weak_b005 --steps 20 --warmup 3 ;;
    weak_b025) run weak_b025 --steps 10 --warmup 3 --beta 0.25 ;;
    weak_b100) run weak_b100 --steps 6 --warmup 3 --beta 1.0 ;;
    strong_b005) run strong_b005 --steps 10 --warmup 3 --scaling strong ;;
  esac
done
python - <<PY
import json, glob
for f in sorted(glob.glob("$OUT/${TAG}_n${N}_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as ex:
        print(f, "unreadable", ex); continue
    print(f.split("/")[-1], "value %.3g edges/s" % d["value"], "ms %.2f" % d["ms_per_step"], "roofline %.3f" % d["roofline"]["frac"],
          "parity", d["parity"]["ok"], d["parity"]["max_rel"], "e2e ms %.1f" % d["e2e"]["ms_per_step"],
          "anchor", d.get("anchor_local_only", {}).get("efficiency_vs_anchor"))
PY
