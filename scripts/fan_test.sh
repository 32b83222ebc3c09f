run() { python bench.py "$@" --no-e2e --no-cpu-baseline --steps 10 --warmup 3 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'])"; }
for bf in 2.5e8 5e8 1e9 2e9; do for n in 1 8; do echo "BIG_FLOPS=$bf shard $n"; SI_G3_BIG_FLOPS=$bf run --shard-of $n; done; done
