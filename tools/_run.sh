B200WOQ_SYRK_PAIR=1 timeout 120 python -m pytest tests/test_kernels_gpu.py -x -q -k "hessian" 2>&1 | tail -5
echo "pair=0"; timeout 120 python tools/bench_kernels.py hessian 2>&1 >/dev/null | grep -o "'C': [0-9]*, 'dtype': 'torch.[a-z0-9]*', 'T': [0-9]*, 'tc_ms': [0-9.]*"
echo "pair=1"; B200WOQ_SYRK_PAIR=1 timeout 120 python tools/bench_kernels.py hessian 2>&1 >/dev/null | grep -o "'C': [0-9]*, 'dtype': 'torch.[a-z0-9]*', 'T': [0-9]*, 'tc_ms': [0-9.]*"
