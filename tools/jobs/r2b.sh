set -x
mkdir -p gpurun_out/r2b
O=gpurun_out/r2b
# page-fault cost on this box
python - > $O/pagefault.log 2>&1 <<'PY'
import mmap, numpy as np, time
sz=512<<20
for hp in (False, True):
    m=mmap.mmap(-1, sz)
    if hp: m.madvise(mmap.MADV_HUGEPAGE)
    a=np.frombuffer(m, dtype=np.uint8)
    t=time.time(); a[:]=1; print('hugepage' if hp else 'normal', sz>>20, 'MB', time.time()-t)
    del a; m.close()
PY
cat $O/pagefault.log
timeout 1700 python -m pytest tests/test_gpu.py -m gpu -x -q > $O/pytest.log 2>&1
tail -15 $O/pytest.log
timeout 600 python bench.py --pairs 2000000 --steps 3 --warmup 3 > $O/bench_2Mpairs.json 2> $O/bench_2Mpairs.err
tail -c 3000 $O/bench_2Mpairs.json; tail -5 $O/bench_2Mpairs.err
