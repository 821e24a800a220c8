set -x
O=gpurun_out/r2i; mkdir -p $O
timeout 1500 python -m pytest tests/test_gpu.py -m gpu -x -q > $O/pytest.log 2>&1
tail -4 $O/pytest.log
python tools/sweep.py synth:4000000 warp_per_read=0 > $O/sweep_4M.log 2>&1; cat $O/sweep_4M.log
timeout 1500 python tools/big_index.py --mbp 512 --out $O/big_index_512Mbp.json > $O/big_index.log 2>&1
tail -30 $O/big_index.log
# seed kernel on the HBM-resident index under ncu (count pass + fill pass)
cat > /tmp/seed_prof.py <<'PY'
import sys, os
sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", "."))
sys.path.insert(0, os.path.join(os.environ.get("GRAFT_REPO_ROOT", "."), "tools"))
import numpy as np, bench, hisat2_b200 as h2
import importlib.util
spec = importlib.util.spec_from_file_location("bi", os.path.join(os.environ.get("GRAFT_REPO_ROOT", "."), "tools", "big_index.py")); bi = importlib.util.module_from_spec(spec); spec.loader.exec_module(bi)
chroms = bi.make_reference(512)
idx = h2.Index(os.path.join(bi.B, "synth512"))
d1, _ = bi.sim(chroms, 1000000 // len(chroms), False)
sb = h2.ReadBatch.parse(data1=d1)
sr = idx.seed_search(sb, max_range=4); print("seed ms", sr.ms_kernel, "lf", sr.n_lf)
PY
ncu --set full --clock-control none -k regex:ht2_seed_kernel -c 2 -o /tmp/seed_big python /tmp/seed_prof.py > $O/ncu_seed.log 2>&1
ncu -i /tmp/seed_big.ncu-rep --page raw --csv > $O/seed_big_raw.csv 2>/dev/null
tail -3 $O/ncu_seed.log; ls -la $O
