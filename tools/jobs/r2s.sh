set -x
O=gpurun_out/r2s; mkdir -p $O
for e in 32 24 16 8 4; do
  echo "== elect $e" >> $O/sweep.log
  HT2GPU_ELECT=$e python tools/sweep.py synth:1000000 warp_per_read=0 >> $O/sweep.log 2>&1
  HT2GPU_ELECT=$e python tools/sweep.py synth:4000000 warp_per_read=0 >> $O/sweep.log 2>&1
done
cat $O/sweep.log
