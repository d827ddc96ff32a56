mkdir -p gpurun_out
python tools/sc_rounds.py 26 1 b3 2>&1 | tail -1 > gpurun_out/r02w_sc26.log
python tools/sc_rounds.py 24 1 e3 2>&1 | tail -1 >> gpurun_out/r02w_sc26.log
cat gpurun_out/r02w_sc26.log
