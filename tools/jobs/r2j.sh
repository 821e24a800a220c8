set -x
O=gpurun_out/r2j; mkdir -p $O
timeout 1500 python -m pytest tests/test_gpu.py -m gpu -x -q > $O/pytest.log 2>&1
tail -4 $O/pytest.log
for bs in 1000000 2000000 4000000; do
  timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras --batch-reads $bs > $O/bench_overlap_$bs.json 2> $O/bench_overlap_$bs.err
  HT2GPU_NO_TAIL_OVERLAP=1 timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras --batch-reads $bs > $O/bench_serial_$bs.json 2> $O/bench_serial_$bs.err
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/r2j/bench_*.json')):
    try:
        d = [json.loads(l) for l in open(f) if l.startswith('{')][-1]
        print(f.split('/')[-1], 'value %.2fM e2e %.2fM  e2e ms/step %.0f kernels %s' % (d['value']/1e6, d['e2e']['value']/1e6, d['e2e']['ms_per_step'], d['kernels_ms_per_step']))
    except Exception as e:
        print(f, 'ERR', e)
PY
tail -3 $O/*.err | tail -20
