set -x
O=gpurun_out/r2c; mkdir -p $O
timeout 600 python bench.py --pairs 2000000 --steps 3 --warmup 3 --no-extras > $O/bench_2Mpairs.json 2> $O/bench_2Mpairs.err
tail -c 2500 $O/bench_2Mpairs.json; tail -5 $O/bench_2Mpairs.err
# launch-configuration sweep of the pool kernel (1M SE synthetic reads, linear index)
python tools/sweep.py synth:1000000 threads_per_block=256,blocks_per_sm=1,slots_per_lane=4 threads_per_block=256,blocks_per_sm=2,slots_per_lane=2 \
   threads_per_block=256,blocks_per_sm=2,slots_per_lane=4 threads_per_block=512,blocks_per_sm=1,slots_per_lane=2 threads_per_block=512,blocks_per_sm=1,slots_per_lane=4 \
   threads_per_block=256,blocks_per_sm=1,slots_per_lane=8 threads_per_block=256,blocks_per_sm=1,slots_per_lane=2 > $O/sweep_cfg.log 2>&1
cat $O/sweep_cfg.log
# ncu: full capture of the alignment kernel and of the SAM kernels (one launch each), source view on
ncu --set full --clock-control none --import-source on -k regex:ht2_align_pool_kernel -c 1 -o $O/pool_1M python tools/prof_run.py synth:1000000 1 > $O/ncu_pool.log 2>&1
ncu --set full --clock-control none -k regex:ht2_sam_kernel -c 2 -o $O/sam_1M python tools/prof_run.py synth:1000000 1 > $O/ncu_sam.log 2>&1
HT2_INDEX=22_20-21M_snp ncu --set full --clock-control none --import-source on -k regex:ht2_align_pool_kernel -c 1 -o $O/pool_graph_200k python tools/prof_run.py synth:200000 1 > $O/ncu_graph.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:ht2_align_pool_kernel -c 1 -o $O/pool_dp2_200k python tools/prof_run.py synth:200000 1 bowtie2_dp=2 > $O/ncu_dp2.log 2>&1
ls -la $O
