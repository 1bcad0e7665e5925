#!/bin/bash
# N GPUs, end state: the 2-GPU peer test, then the driver's bench command
N=${1:-2}
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_render.py -m gpu -x -q -k "peer" > gpurun_out/m${N}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/m${N}_pytest.log
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 100 --warmup 10 > gpurun_out/m${N}_bench.json 2> gpurun_out/m${N}_bench.err
echo "bench rc=$?"
python - "gpurun_out/m${N}_bench.json" <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("%.4f ms" % d["ms_per_step"], "%.2f M rays/s" % (d["value"] / 1e6), "sustained %.2f M" % (d["sustained"]["value"] / 1e6), "e2e %.2f M" % (d["e2e"]["value"] / 1e6), d.get("exchange"))
for k in ("strong_scaling", "full_frame", "tracking", "single_frame"):
    print(k, json.dumps(d.get(k))[:300])
PY
