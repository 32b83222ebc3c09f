run() { python bench.py "$@" --no-e2e --no-cpu-baseline --steps 10 --warmup 3 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'])"; }
for w in 4 2 1; do for n in 1 8; do echo "WARPS_PER_BLOCK=$w shard $n"; SI_G3_WARPS_PER_BLOCK=$w run --shard-of $n; done; done
