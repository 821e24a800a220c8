set -x
O=gpurun_out/r2y; mkdir -p $O
timeout 1200 python -m pytest tests/test_gpu.py -m gpu -x -q -k "populated_splice or spliced_alignment or command_line or pipeline_reads" > $O/pytest_ss.log 2>&1; tail -15 $O/pytest_ss.log
