set -x
O=gpurun_out/r2v; mkdir -p $O
HT2_INDEX=22_20-21M_snp python tools/sweep.py synth:500000 warp_per_read=0 > $O/sweep_graph.log 2>&1; cat $O/sweep_graph.log
python tools/sweep.py synth:4000000 warp_per_read=0 > $O/sweep_linear.log 2>&1; cat $O/sweep_linear.log
timeout 2400 python -m pytest tests/test_gpu.py -m gpu -x -q > $O/pytest.log 2>&1; tail -5 $O/pytest.log
