mkdir -p gpurun_out
python -m pytest tests/test_gpu_basefold_sharded.py tests/test_gpu_basefold.py -m gpu -x -q 2>&1 | tail -25 > gpurun_out/r02r_shard.log
cat gpurun_out/r02r_shard.log
