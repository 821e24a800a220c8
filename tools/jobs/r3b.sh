set -x
O=gpurun_out/r3b; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu.py -m gpu -x -q -k "two_devices or torchrun" > $O/pytest_2gpu.log 2>&1; tail -3 $O/pytest_2gpu.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > $O/bench_2gpu.json 2> $O/bench_2gpu.err; tail -c 700 $O/bench_2gpu.json; tail -3 $O/bench_2gpu.err
