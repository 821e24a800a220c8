set -x
O=gpurun_out/r2n; mkdir -p $O
python -c "from hisat2_b200 import api; print(api.sw_selftest(3000, 5))" > $O/selftest.log 2>&1; cat $O/selftest.log
timeout 1500 python -m pytest tests/test_gpu.py -m gpu -x -q -k "dynamic_programming or warp_wide or option_matrix" > $O/pytest_dp.log 2>&1; tail -5 $O/pytest_dp.log
python tools/sweep.py synth:200000 bowtie2_dp=0 bowtie2_dp=1 bowtie2_dp=2 > $O/sweep_dp.log 2>&1; cat $O/sweep_dp.log
HT2GPU_STATS=1 python tools/sweep.py synth:200000 bowtie2_dp=2 2>&1 | tail -48 | cut -c1-110 > $O/stats_dp2.log
head -50 $O/stats_dp2.log
timeout 2400 python -m pytest tests/test_gpu.py -m gpu -x -q > $O/pytest.log 2>&1; tail -5 $O/pytest.log
