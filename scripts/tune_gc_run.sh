#!/bin/bash
# on the GPU box: per-class coul timings for every library variant
V=sympleints_b200/lib/variants_gc
for f in $V/lib_sh*_mb*.so; do
  tag=$(basename $f .so); tag=${tag#lib_}
  SYMPLEINTS_B200_LIB=$PWD/$f SI_GC_PROFILE=1 python scripts/coul_bench.py 2 > /dev/null 2> gpurun_out/tunegc_$tag.log
  echo $tag $(grep -c GCPROF gpurun_out/tunegc_$tag.log)
done
