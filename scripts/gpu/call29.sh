set -x
O=gpurun_out/r02; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_tc_engine.py tests/test_gpu_properties.py -x -q -m gpu 2>&1 | tail -3
timeout 300 python scripts/time_kernels.py contract --shapes mesh1,mesh2 --reps 20 2>&1 | tee $O/time_contract_b.txt
timeout 300 python scripts/time_kernels.py spmm --shapes mesh1,mesh2,rgg --rowtile 4 --rtmodes 2,3 --reps 10 2>&1 | tee $O/spmm_modes_small_blocks.txt
timeout 600 python scripts/sweep.py --train --quick 2>&1 | tee $O/sweep_train_n1.csv | tail -25
