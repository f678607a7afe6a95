#!/bin/bash
# round 2, 2-GPU call: sharded proof over peer windows (log 16 parity vs single GPU, log 20 timing, NCCL path A/B)
cd "$(dirname "$0")/.."
run() { timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $1 tests/sharded_proof_worker.py $2 single 3 2>&1 | grep -E "SHARDED_OK|Error|error|assert" | head -5; }
echo "== p2p log16"; run 29541 16
echo "== p2p log20"; run 29542 20
echo "== nccl log20"; S2C_NO_P2P=1 run 29543 20
echo "== p2p log18"; run 29544 18
