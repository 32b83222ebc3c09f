#!/bin/bash
set -e
mkdir -p sympleints_b200/lib/variants
rm -f sympleints_b200/lib/variants/*.so
for skip in boys,vrr,dtree,hrr,c2sa prim,hrr,c2sa; do
  tag=$(echo $skip | tr ',' '_')
  SI_G3_SKIP=$skip python -c "from sympleints_b200 import build; build.build_library(force=True)" > /dev/null
  cp sympleints_b200/lib/libsympleints_b200.so sympleints_b200/lib/variants/lib_abl_$tag.so
  echo built $tag
done
python -c "from sympleints_b200 import build; build.build_library(force=True)" > /dev/null
