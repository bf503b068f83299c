timeout 300 python -m pytest tests/test_kernels_gpu.py -x -q -k "hessian" 2>&1 | tail -3
for sr in 1 4 8 16 32; do echo "super_rows=$sr"; B200WOQ_SYRK_SUPER_ROWS=$sr timeout 300 python tools/bench_kernels.py hessian 2>&1 >/dev/null | grep -o "'C': [0-9]*\|float16\|bfloat16\|'tc_ms': [0-9.]*" | paste - - - ; done
