#!/bin/bash
cd "$(dirname "$0")/.."
run() { timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $1 profiles/sharded_stages.py $2 3 2>&1 | grep -E '^\{|Error|error' | head -6; }
echo "== p2p log20"; run 29542 20 | tee gpurun_out/r02h2_p2p.jsonl
echo "== nccl log20"; S2C_NO_P2P=1 run 29543 20 | tee gpurun_out/r02h2_nccl.jsonl
