set -x
O=gpurun_out/r2f; mkdir -p $O
timeout 1500 python -m pytest tests/test_gpu.py -m gpu -x -q > $O/pytest.log 2>&1
tail -5 $O/pytest.log
python tools/sweep.py synth:1000000 warp_per_read=0 > $O/sweep_1M.log 2>&1
python tools/sweep.py synth:4000000 warp_per_read=0 threads_per_block=128,slots_per_lane=8 threads_per_block=128,slots_per_lane=4 threads_per_block=128,slots_per_lane=8,blocks_per_sm=2 > $O/sweep_4M.log 2>&1
cat $O/sweep_*.log
bash tools/jobs/prof_all.sh $O
ls -la $O; du -sh $O
