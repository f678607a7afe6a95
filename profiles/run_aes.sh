python profiles/aes_bench.py 16 8 3
python profiles/aes_bench.py 16 12 3
python profiles/aes_bench.py 16 16 3
python profiles/aes_bench.py 32 16 2
