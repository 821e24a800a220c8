#!/bin/bash
# oracle/make_data.sh -- stage the reference's example DATA (index, FASTA,
# reads: data files, not sources) plus generated inputs under
# oracle/_ref/data/ (git-ignored, shipped to the GPU box by gpurun), build the
# linear 22_20-21M index with the reference's own hisat2-build-s, and produce
# reference SAMs with the unmodified reference binary.  Test infrastructure.
set -e
cd "$(dirname "$0")"
REF=${REF:-/root/reference}
D=_ref/data
mkdir -p $D
cp -n $REF/example/index/22_20-21M_snp.*.ht2 $D/
cp -n $REF/example/reference/22_20-21M.fa $REF/example/reference/22_20-21M.snp $D/
cp -n $REF/example/reads/reads_1.fa $REF/example/reads/reads_2.fa $D/
if [ ! -f $D/22_20-21M.1.ht2 ]; then
  ./_ref/hisat2-build-s -q $D/22_20-21M.fa $D/22_20-21M > /dev/null
fi
if [ ! -f $D/sim10k_1.fa ]; then
  python ../tools/simreads.py $D/22_20-21M.fa 10000 $D/sim10k --seed 1 --paired
fi
