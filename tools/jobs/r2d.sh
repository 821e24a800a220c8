set -x
O=gpurun_out/r2d; mkdir -p $O
timeout 1500 python -m pytest tests/test_gpu.py -m gpu -x -q > $O/pytest.log 2>&1
tail -5 $O/pytest.log
python tools/sweep.py synth:200000 bowtie2_dp=0 bowtie2_dp=1 bowtie2_dp=2 > $O/sweep_dp_200k.log 2>&1
python tools/sweep.py synth:1000000 bowtie2_dp=0 bowtie2_dp=2 > $O/sweep_dp_1M.log 2>&1
python tools/sweep.py synth:2000000 warp_per_read=0 > $O/sweep_2M.log 2>&1
python tools/sweep.py synth:4000000 warp_per_read=0 > $O/sweep_4M.log 2>&1
HT2_INDEX=22_20-21M_snp python tools/sweep.py synth:1000000 warp_per_read=0 > $O/sweep_graph_1M.log 2>&1
cat $O/sweep_*.log
HT2GPU_STATS=1 HT2_INDEX=22_20-21M_snp python tools/sweep.py synth:1000000 warp_per_read=0 > $O/stats_graph_1M.log 2>&1
# lean ncu: keep CSV pages, not the reports
ncu --set full --clock-control none --import-source on -k regex:ht2_align_pool_kernel -c 1 -o /tmp/pool_1M python tools/prof_run.py synth:1000000 1 > $O/ncu_pool.log 2>&1
ncu -i /tmp/pool_1M.ncu-rep --page raw --csv > $O/pool_1M_raw.csv 2>/dev/null
ncu -i /tmp/pool_1M.ncu-rep --page source --csv 2>/dev/null | gzip -9 > $O/pool_1M_source.csv.gz
ncu --set full --clock-control none -k regex:ht2_sam_kernel -c 2 -o /tmp/sam_1M python tools/prof_run.py synth:1000000 1 > $O/ncu_sam.log 2>&1
ncu -i /tmp/sam_1M.ncu-rep --page raw --csv > $O/sam_1M_raw.csv 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_bench.csv python bench.py --pairs 2000000 --steps 2 --warmup 3 --no-extras --no-cpu-baseline > $O/bench_under_ncu.log 2>&1
ls -la $O; du -sh $O
