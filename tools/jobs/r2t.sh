set -x
O=gpurun_out/r2t; mkdir -p $O
export HT2_INDEX=22_20-21M_snp
for e in 32 8 2; do
  echo "== graph elect $e" >> $O/sweep.log
  HT2GPU_ELECT=$e python tools/sweep.py synth:500000 warp_per_read=0 threads_per_block=128,slots_per_lane=8 threads_per_block=128,slots_per_lane=4 >> $O/sweep.log 2>&1
done
cat $O/sweep.log
