#!/bin/bash
cd "$(dirname "$0")/.."
run() { timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $1 tests/sharded_proof_worker.py $2 single 3 2>&1 | grep -E "SHARDED_OK|Error|error|assert" | head -5; }
prof() { timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $1 profiles/sharded_stages.py $2 3 2>&1 | grep -E '^\{|Error|error' | head -6; }
echo "== p2p log16"; run 29541 16
echo "== nccl log16"; S2C_NO_P2P=1 run 29545 16
echo "== p2p log20"; run 29542 20
echo "== p2p log20 stages"; prof 29546 20 | tee gpurun_out/r02h3_p2p.jsonl
