set -x
O=gpurun_out/r2u; mkdir -p $O
timeout 1700 python tools/big_index.py --mbp 1024 --threads 96 --out $O/big_index_1024Mbp.json > $O/big.log 2>&1
tail -5 $O/big.log
