#!/bin/bash
# round 2, 8-GPU call: sharded proof over peer windows at log 20 (stages), log 22, log 24 (reference verifier)
cd "$(dirname "$0")/.."
G=${G:-8}
prof() { timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port $1 profiles/sharded_stages.py $2 3 2>&1 | grep -E '^\{|Error|error' | head -8; }
run() { timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port $1 tests/sharded_proof_worker.py $2 $3 $4 2>&1 | grep -E "SHARDED_OK|Error|error|assert" | head -5; }
echo "== log20 single-GPU parity"; run 29541 20 single 3 | tee gpurun_out/r02i8_log20_parity.txt
echo "== log20 stages"; prof 29546 20 | tee gpurun_out/r02i8_log20.jsonl
echo "== log24 ref"; run 29547 24 ref 2 | tee gpurun_out/r02i8_log24.txt
