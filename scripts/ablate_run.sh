#!/bin/bash
for f in sympleints_b200/lib/variants/lib_abl_*.so; do
  tag=$(basename $f .so); tag=${tag#lib_abl_}
  SYMPLEINTS_B200_LIB=$PWD/$f SI_G3_PROFILE=1 python bench.py --steps 2 --warmup 0 --no-cpu-baseline --no-e2e > /dev/null 2> gpurun_out/abl_$tag.log
  echo $tag $(grep -c G3PROF gpurun_out/abl_$tag.log)
done
