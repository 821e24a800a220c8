set -x
O=gpurun_out/r2q; mkdir -p $O
cd /tmp && rm -rf cubx && mkdir cubx && cd cubx && cuobjdump -xelf all $GRAFT_REPO_ROOT/hisat2_b200/libht2gpu.so > /dev/null 2>&1 && nvdisasm -c ht2_gpu.sm_100a.cubin > all.sass 2>/dev/null; cd $GRAFT_REPO_ROOT
name=pool_dp2_200k
HT2_INDEX=22_20-21M ncu --set full --clock-control none --import-source on -k regex:ht2_align_pool_kernel -c 1 -o /tmp/$name python tools/prof_run.py synth:200000 1 bowtie2_dp=2 > $O/ncu_$name.log 2>&1
ncu -i /tmp/$name.ncu-rep --page raw --csv > $O/${name}_raw.csv 2>/dev/null
ncu -i /tmp/$name.ncu-rep --page source --csv 2>/dev/null | gzip -9 > /tmp/${name}_source.csv.gz
NCU_FUNC_DETAIL=swFillCoop python tools/ncu_funcs.py /tmp/${name}_source.csv.gz /tmp/cubx/all.sass ht2_align_pool_kernelILi8ELi4ELb0ELb1E $O/${name}_functions.json > $O/${name}_functions.txt 2>&1
python tools/ncu_summary.py /tmp/$name.ncu-rep $O/${name}_summary.json "$name" > /dev/null 2>&1
head -80 $O/${name}_functions.txt
