python -m pytest tests -m gpu -x -q -k "aes or lde_packed or pool" 2>&1 | tail -5
python profiles/aes_bench.py 16 8 3
python profiles/aes_bench.py 16 16 3
python profiles/aes_bench.py 32 16 3
