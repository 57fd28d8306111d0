timeout 600 python -m pytest tests/test_gpu_tc_engine.py tests/test_gpu_fullsize.py -x -q -m gpu 2>&1 | tail -2
for f in "" 1; do echo "FUSEA=$f"; TGCN_T3_FUSEA=$f timeout 300 python scripts/time_kernels.py contract --shapes rgg --reps 5 2>&1 | grep bwd_w; done
