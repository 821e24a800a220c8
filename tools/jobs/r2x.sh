set -x
O=gpurun_out/r2z; mkdir -p $O
timeout 2400 python -m pytest tests/test_gpu.py -m gpu -x -q > $O/pytest.log 2>&1; tail -4 $O/pytest.log
# profile of the final kernel on a bench-shaped launch (2 M pairs = 4 M reads)
cd /tmp && rm -rf cubx && mkdir cubx && cd cubx && cuobjdump -xelf all $GRAFT_REPO_ROOT/hisat2_b200/libht2gpu.so > /dev/null 2>&1 && nvdisasm -c ht2_gpu.sm_100a.cubin > all.sass 2>/dev/null; cd $GRAFT_REPO_ROOT
ncu --set full --clock-control none --import-source on -k regex:ht2_align_pool_kernel -c 1 -o /tmp/pool_pe python tools/prof_run.py synthpe:2000000 1 no_spliced_alignment=1 > $O/ncu_pool_pe.log 2>&1
ncu -i /tmp/pool_pe.ncu-rep --page raw --csv > $O/pool_pe_4M_raw.csv 2>/dev/null
ncu -i /tmp/pool_pe.ncu-rep --page source --csv 2>/dev/null | gzip -9 > /tmp/pool_pe_source.csv.gz
python tools/ncu_funcs.py /tmp/pool_pe_source.csv.gz /tmp/cubx/all.sass ht2_align_pool_kernelILi8ELi4ELb0ELb1E $O/pool_pe_4M_functions.json > $O/pool_pe_4M_functions.txt 2>&1
python tools/ncu_summary.py /tmp/pool_pe.ncu-rep $O/pool_pe_4M_summary.json "final kernel, 2 M pairs (4 M reads) per launch" > /dev/null 2>&1
python tools/make_traffic.py $O/pool_pe_4M_summary.json profiles/traffic.json "ncu --set full, one launch of ht2_align_pool_kernel<8,4,false,true> on 2 M synthetic pairs (the bench's launch shape), profiles/r02_ncu_pool_pe_4M_summary.json" > $O/traffic.log 2>&1
cp profiles/traffic.json $O/traffic.json
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > $O/bench_final.json 2> $O/bench_final.err
tail -c 300 $O/bench_final.json; tail -3 $O/bench_final.err
timeout 600 python bench.py --impl reference --gpus 1 --steps 5 --warmup 2 > $O/bench_reference.json 2> $O/bench_reference.err
tail -c 400 $O/bench_reference.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_bench.csv python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline > $O/bench_under_ncu.log 2>&1
ls -la $O
