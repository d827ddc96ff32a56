mkdir -p gpurun_out
python -m pytest tests/test_gpu_matmul.py tests/test_zkml.py tests/test_gpu_witness.py -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r03e_matmul.log
cat gpurun_out/r03e_matmul.log
