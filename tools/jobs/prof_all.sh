# usage: prof_all.sh <outdir>   -- ncu captures reduced to small files ON THE BOX (reports stay in /tmp)
O=$1; mkdir -p $O
cd /tmp && rm -rf cubx && mkdir cubx && cd cubx && cuobjdump -xelf all $GRAFT_REPO_ROOT/hisat2_b200/libht2gpu.so > /dev/null 2>&1 && nvdisasm -c ht2_gpu.sm_100a.cubin > all.sass 2>/dev/null; cd $GRAFT_REPO_ROOT
prof() {  # name kernel-substring env-index reads extra-opts...
  local name=$1 ksub=$2 index=$3 reads=$4; shift 4
  HT2_INDEX=$index ncu --set full --clock-control none --import-source on -k regex:ht2_align_pool_kernel -c 1 -o /tmp/$name python tools/prof_run.py $reads 1 "$@" > $O/ncu_$name.log 2>&1
  ncu -i /tmp/$name.ncu-rep --page raw --csv > $O/${name}_raw.csv 2>/dev/null
  ncu -i /tmp/$name.ncu-rep --page source --csv 2>/dev/null | gzip -9 > /tmp/${name}_source.csv.gz
  python tools/ncu_funcs.py /tmp/${name}_source.csv.gz /tmp/cubx/all.sass $ksub $O/${name}_functions.json > $O/${name}_functions.txt 2>&1
  python tools/ncu_summary.py /tmp/$name.ncu-rep $O/${name}_summary.json "$name" > /dev/null 2>&1
}
prof pool_linear_1M ht2_align_pool_kernelILi8ELi4ELb0ELb1E 22_20-21M synth:1000000
prof pool_graph_500k ht2_align_pool_kernelILi8ELi4ELb1ELb1E 22_20-21M_snp synth:500000
prof pool_dp2_200k ht2_align_pool_kernelILi8ELi4ELb0ELb1E 22_20-21M synth:200000 bowtie2_dp=2
ncu --set full --clock-control none -k regex:ht2_sam_kernel -c 2 -o /tmp/sam_1M python tools/prof_run.py synth:1000000 1 > $O/ncu_sam.log 2>&1
ncu -i /tmp/sam_1M.ncu-rep --page raw --csv > $O/sam_1M_raw.csv 2>/dev/null
