#!/bin/bash
# on the GPU box: per-class timings for every library variant
for f in sympleints_b200/lib/variants/lib_sh*_mb*.so; do
  tag=$(basename $f .so); tag=${tag#lib_}
  SYMPLEINTS_B200_LIB=$PWD/$f SI_G3_PROFILE=1 python bench.py --steps 2 --warmup 0 --no-cpu-baseline --no-e2e --no-other-keys --no-parity > /dev/null 2> gpurun_out/tune_$tag.log
  echo $tag $(grep -c G3PROF gpurun_out/tune_$tag.log)
done
