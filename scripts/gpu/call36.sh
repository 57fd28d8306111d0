for ru in 32 16; do echo "RU=$ru"; TGCN_T3_RU=$ru timeout 300 python scripts/time_kernels.py contract --shapes mesh1,mesh2 --reps 10 2>&1 | grep bwd_w; done
