set -x
O=gpurun_out/r2r; mkdir -p $O
timeout 1500 python -m pytest tests/test_gpu.py -m gpu -x -q -k "dynamic_programming or warp_wide or option_matrix" > $O/pytest_dp.log 2>&1; tail -3 $O/pytest_dp.log
python tools/sweep.py synth:200000 bowtie2_dp=1 bowtie2_dp=2 > $O/sweep_dp.log 2>&1; cat $O/sweep_dp.log
HT2GPU_STATS=1 python tools/sweep.py synth:200000 bowtie2_dp=2 2>&1 | grep "T_HYB_DP\|F_ENTER\|T_PS " | tail -3 | cut -c1-160
