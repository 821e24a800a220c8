set -x
O=gpurun_out/r2g; mkdir -p $O
timeout 1500 python -m pytest tests/test_gpu.py -m gpu -x -q --durations=8 > $O/pytest.log 2>&1
tail -14 $O/pytest.log
python tools/sweep.py synth:200000 bowtie2_dp=0 bowtie2_dp=1 bowtie2_dp=2 > $O/sweep_dp_200k.log 2>&1
HT2_INDEX=22_20-21M_snp python tools/sweep.py synth:1000000 warp_per_read=0 > $O/sweep_graph_1M.log 2>&1
python tools/sweep.py synth:4000000 warp_per_read=0 > $O/sweep_4M.log 2>&1
cat $O/sweep_*.log
timeout 900 python bench.py --steps 3 --warmup 3 > $O/bench_10Mpairs.json 2> $O/bench_10Mpairs.err
tail -c 600 $O/bench_10Mpairs.json; tail -5 $O/bench_10Mpairs.err
