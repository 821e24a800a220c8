set -x
O=gpurun_out/r3a; mkdir -p $O
timeout 2400 python -m pytest tests/test_gpu.py -m gpu -x -q > $O/pytest.log 2>&1; tail -4 $O/pytest.log
