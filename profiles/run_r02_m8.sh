#!/bin/bash
# round 2, 8-GPU call: the driver's N=8 launch line (weak scaling + sharded record), sharded stage profile, log 24 with the reference verifier
cd "$(dirname "$0")/.."
G=${G:-8}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $G --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02m${G}_bench.json 2> gpurun_out/r02m${G}_bench.err
tail -2 gpurun_out/r02m${G}_bench.err
python - <<PY
import json
for l in open("gpurun_out/r02m${G}_bench.json"):
    if l.startswith("{"):
        d = json.loads(l)
        print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["ms_per_step"])
        print(json.dumps(d.get("sharded"))[:1800])
        print(json.dumps(d.get("proof_batches_cfg4"))[:700])
PY
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29546 profiles/sharded_stages.py 20 3 2>&1 | grep -E '^\{|Error|error' | head -8 | tee gpurun_out/r02m${G}_log20.jsonl | cut -c1-900
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29547 tests/sharded_proof_worker.py 24 ref 2 2>&1 | grep -E "SHARDED_OK|Error|error|assert" | head -5 | tee gpurun_out/r02m${G}_log24.txt
