python profiles/small_proofs_bench.py chacha20 256
python profiles/small_proofs_bench.py aes128 128
nproc
