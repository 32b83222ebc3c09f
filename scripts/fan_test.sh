run() { python bench.py "$@" --no-e2e --no-cpu-baseline --steps 10 --warmup 3 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['gpu_launches'])"; }
for m in 1 0; do for n in 1 2 4 8; do echo "MERGE=$m shard $n"; SI_G3_MERGE=$m run --shard-of $n; done; done
