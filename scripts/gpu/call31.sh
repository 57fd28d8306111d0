set -x
O=gpurun_out/r02; mkdir -p $O
for ru in 32 16 8; do echo "RU=$ru"; TGCN_T3_RU=$ru timeout 300 python scripts/time_kernels.py contract --shapes mesh1,mesh2 --reps 10 2>&1 | grep bwd_w; done | tee $O/bwd_w_ru.txt
TGCN_T3_RU=16 timeout 300 python -m pytest tests/test_gpu_tc_engine.py -x -q -m gpu 2>&1 | tail -2
timeout 600 python scripts/sweep.py --train --quick 2>&1 | grep -v "^+" | tee $O/sweep_train_n1.csv | tail -20
