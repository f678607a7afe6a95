python -m pytest tests -m gpu -x -q -k "lde_packed" 2>&1 | tail -3
python profiles/fft_bench.py 16 16 2
python profiles/fft_bench.py 20 16 3
