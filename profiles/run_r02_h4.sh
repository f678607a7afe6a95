#!/bin/bash
cd "$(dirname "$0")/.."
G=${G:-2}
run() { timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port $1 tests/sharded_proof_worker.py $2 single 4 2>&1 | grep -E "SHARDED_OK|Error|error|assert" | head -5; }
echo "== 2-stream log16"; run 29541 16
echo "== 2-stream log18"; run 29542 18
echo "== 2-stream log20"; run 29543 20
echo "== 1-stream log20"; S2C_P2P_1STREAM=1 run 29544 20
