set -x
O=gpurun_out/r2e; mkdir -p $O
timeout 1500 python -m pytest tests/test_gpu.py -m gpu -x -q > $O/pytest.log 2>&1
tail -5 $O/pytest.log
python tools/sweep.py synth:1000000 warp_per_read=0 > $O/sweep_1M.log 2>&1
python tools/sweep.py synth:4000000 warp_per_read=0 slots_per_lane=2 > $O/sweep_4M.log 2>&1
python tools/sweep.py synth:200000 bowtie2_dp=0 bowtie2_dp=2 > $O/sweep_dp_200k.log 2>&1
HT2_INDEX=22_20-21M_snp python tools/sweep.py synth:1000000 warp_per_read=0 > $O/sweep_graph_1M.log 2>&1
cat $O/sweep_*.log
HT2GPU_STATS=1 python tools/sweep.py synth:1000000 warp_per_read=0 2>&1 | tail -45 > $O/stats_linear_1M.log
timeout 900 python bench.py --steps 3 --warmup 3 > $O/bench_10Mpairs.json 2> $O/bench_10Mpairs.err
tail -c 1500 $O/bench_10Mpairs.json; tail -5 $O/bench_10Mpairs.err
ls -la $O; du -sh $O
