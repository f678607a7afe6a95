#!/bin/bash
# round 2, 2-GPU call: the driver's N=2 launch line (sharded record inside) + the 2-rank sharded parity tests
cd "$(dirname "$0")/.."
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02g2_bench.json 2> gpurun_out/r02g2_bench.err
tail -3 gpurun_out/r02g2_bench.err
python - <<'PY'
import json
for l in open("gpurun_out/r02g2_bench.json"):
    if l.startswith("{"):
        d = json.loads(l)
        print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["ms_per_step"])
        print(json.dumps(d.get("sharded"))[:1500])
        print(json.dumps(d.get("proof_batches_cfg4"))[:600])
PY
timeout 600 python -m pytest tests/test_gpu_sharded.py -x -q -m gpu 2>&1 | tail -3
