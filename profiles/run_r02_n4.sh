#!/bin/bash
cd "$(dirname "$0")/.."
G=${G:-4}
run() { timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port $1 tests/sharded_proof_worker.py $2 single 4 2>&1 | grep -E "SHARDED_OK|Error|error|assert" | head -5; }
echo "== occ1 log20"; S2C_P2P_OCC=1 run 29543 20
echo "== occ2 log20"; S2C_P2P_OCC=2 run 29544 20
echo "== occ3 log20"; S2C_P2P_OCC=3 run 29545 20
