set -x
O=gpurun_out/r2m; mkdir -p $O
for v in "" _noslice _slice32; do
  export HT2GPU_LIB=$PWD/hisat2_b200/libht2gpu$v.so
  echo "== lib $v" >> $O/sweep.log
  python tools/sweep.py synth:1000000 warp_per_read=0 >> $O/sweep.log 2>&1
  python tools/sweep.py synth:4000000 warp_per_read=0 >> $O/sweep.log 2>&1
  HT2_INDEX=22_20-21M_snp python tools/sweep.py synth:1000000 warp_per_read=0 >> $O/sweep.log 2>&1
done
cat $O/sweep.log
unset HT2GPU_LIB
HT2GPU_STATS=1 python tools/sweep.py synth:200000 bowtie2_dp=2 2>&1 | tail -48 | cut -c1-110 > $O/stats_dp2.log
cat $O/stats_dp2.log | head -50
