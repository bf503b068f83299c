timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
B200WOQ_BENCH_SHAPES=4096x4096,11008x4096,4096x11008,12288x4096,22016x4096 timeout 300 python tools/bench_kernels.py gemvs 2>&1 >/dev/null | grep -o "'shape': '[a-z/0-9]*\|'M': [0-9]*\|'us': [0-9.]*" | paste - - -
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step']); print(json.dumps(d['decode']['variants'], indent=1))"
