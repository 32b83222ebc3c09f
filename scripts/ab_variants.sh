# A/B of library variants under sympleints_b200/lib/variants/lib_<tag>.so against the in-tree build ("base")
for v in base ${AB_TAGS:-prev} base ${AB_TAGS:-prev}; do
  if [ $v = base ]; then unset SYMPLEINTS_B200_LIB; else export SYMPLEINTS_B200_LIB=$PWD/sympleints_b200/lib/variants/lib_$v.so; fi
  timeout 300 python bench.py --no-e2e --no-cpu-baseline --no-other-keys --steps 8 > gpurun_out/ab_$v.json 2>/dev/null
  python -c "
import json; d=json.load(open('gpurun_out/ab_$v.json')); print('$v', d['ms_per_step'], d['parity']['failed'])"
done
