set -x
O=gpurun_out/r2l; mkdir -p $O
timeout 1500 python -m pytest tests/test_gpu.py -m gpu -x -q > $O/pytest.log 2>&1
tail -6 $O/pytest.log
python tools/sweep.py synth:1000000 warp_per_read=0 > $O/sweep_1M.log 2>&1
python tools/sweep.py synth:4000000 warp_per_read=0 > $O/sweep_4M.log 2>&1
python tools/sweep.py synth:200000 bowtie2_dp=0 bowtie2_dp=1 bowtie2_dp=2 > $O/sweep_dp_200k.log 2>&1
HT2_INDEX=22_20-21M_snp python tools/sweep.py synth:1000000 warp_per_read=0 > $O/sweep_graph_1M.log 2>&1
cat $O/sweep_*.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $O/bench.json 2> $O/bench.err
tail -c 300 $O/bench.json; tail -3 $O/bench.err
