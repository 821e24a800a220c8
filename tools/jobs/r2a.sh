set -x
mkdir -p gpurun_out/r2a
nproc > gpurun_out/r2a/nproc.txt; free -g >> gpurun_out/r2a/nproc.txt
# spliced-build parity on the GPU
HT2GPU_LIB=$PWD/hisat2_b200/libht2gpu_spl.so timeout 900 python -m pytest tests/test_gpu.py -m gpu -x -q -k "spliced or tiny_se or tiny_pe or chr22_matches or chr22_paired" > gpurun_out/r2a/pytest_spliced.log 2>&1
tail -5 gpurun_out/r2a/pytest_spliced.log
# perf: default vs spliced build, linear SE 1M
python tools/sweep.py synth:1000000 warp_per_read=0 > gpurun_out/r2a/sweep_default.log 2>&1
HT2GPU_LIB=$PWD/hisat2_b200/libht2gpu_spl.so python tools/sweep.py synth:1000000 warp_per_read=0 > gpurun_out/r2a/sweep_spl.log 2>&1
# dp modes + graph: baseline perf numbers
python tools/sweep.py synth:200000 bowtie2_dp=0 bowtie2_dp=1 bowtie2_dp=2 > gpurun_out/r2a/sweep_dp.log 2>&1
HT2_INDEX=22_20-21M_snp python tools/sweep.py synth:200000 warp_per_read=0 > gpurun_out/r2a/sweep_graph.log 2>&1
HT2GPU_STATS=1 python tools/sweep.py synth:1000000 warp_per_read=0 > gpurun_out/r2a/stats_linear.log 2>&1
HT2GPU_STATS=1 HT2_INDEX=22_20-21M_snp python tools/sweep.py synth:200000 warp_per_read=0 > gpurun_out/r2a/stats_graph.log 2>&1
cat gpurun_out/r2a/sweep_*.log
