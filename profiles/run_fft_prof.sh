python -m pytest tests -m gpu -x -q -k "lde_packed" 2>&1 | tail -5
python profiles/fft_bench.py 16 16 2
python profiles/fft_bench.py 18 16 2
python profiles/fft_bench.py 20 16 2
