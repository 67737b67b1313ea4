#!/bin/sh
# On an N-GPU box: the multi-device tests, then bench.py at 8 / 4 / 2 ranks (e2e + concurrent H2D ceiling per N), into gpurun_out/.
# usage: gpurun --gpus 8 -- sh tools/scale_run.sh [tag]
TAG=${1:-r2}
nvidia-smi topo -m > gpurun_out/${TAG}_topo.txt 2>&1
python -m pytest tests -m gpu -q -k "multi_device or c_batch or golden or float or ragged" 2>&1 | tail -6 > gpurun_out/${TAG}_pytest_multi.txt
tail -3 gpurun_out/${TAG}_pytest_multi.txt
for N in 8 4 2; do
  EXTRA="--no-also"; [ "$N" = 8 ] && EXTRA=""
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 $EXTRA \
    > gpurun_out/${TAG}_bench_n$N.json 2> gpurun_out/${TAG}_bench_n$N.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/${TAG}_bench_n$N.json").read().strip().splitlines()[-1])
    e = d["e2e"]
    print("N=$N value %.2f M clips/s  e2e %.2f M (%.1f GB/s, ceiling %.1f GB/s, frac %.3f)" % (d["value"] / 1e6, e["value"] / 1e6, e["gbs"], e["ceiling_gbs"], e["frac_of_ceiling"]))
    for a in d.get("also", []):
        print("   also:", a["config"]["workload"][:40], "%.2f M clips/s" % (a["value"] / 1e6), "frac %.3f" % a["roofline"]["frac"])
except Exception as ex:
    print("N=$N failed:", ex)
PY
done
