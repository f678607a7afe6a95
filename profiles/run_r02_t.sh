#!/bin/bash
cd "$(dirname "$0")/.."
python profiles/small_latency.py chacha20 2 | cut -c1-330
S2C_LEAVES_SEQ_MIX=1 python profiles/small_latency.py chacha20 2 | cut -c1-330
python profiles/small_latency.py chacha20 1024 | cut -c1-330
S2C_LEAVES_SEQ_MIX=1 python profiles/small_latency.py chacha20 1024 | cut -c1-330
S2C_LEAVES_SEQ_MIX=1 timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "golden or fixture" 2>&1 | tail -2
