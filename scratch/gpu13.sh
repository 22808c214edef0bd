VBK_SEGV_TRACE=1 timeout 900 python -m pytest tests/test_gpu.py -m gpu -x -q -s -k "strict_mode_big or batch" > gpurun_out/s3_pytest4.log 2>&1; grep -v "^  File" gpurun_out/s3_pytest4.log | head -80
