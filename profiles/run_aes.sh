python -m pytest tests -m gpu -x -q -k "aes" 2>&1 | tail -25
