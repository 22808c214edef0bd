timeout 900 python -m pytest tests/test_gpu.py -m gpu -x -q -k "sparse_columns_bit_exact or multicommodity" 2>&1 | tail -8
