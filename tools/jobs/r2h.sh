set -x
O=gpurun_out/r2h; mkdir -p $O
nvidia-smi -L > $O/gpus.txt
timeout 1500 python -m pytest tests/test_gpu.py -m gpu -x -q --durations=5 > $O/pytest.log 2>&1
tail -12 $O/pytest.log
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras > $O/bench_n1.json 2> $O/bench_n1.err
tail -c 400 $O/bench_n1.json; tail -3 $O/bench_n1.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > $O/bench_n2.json 2> $O/bench_n2.err
tail -c 400 $O/bench_n2.json; tail -3 $O/bench_n2.err
