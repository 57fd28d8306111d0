set -x
O=gpurun_out/r02; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_multi.py -x -q -m gpu 2>&1 | tail -15
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 scripts/sweep.py --train --quick 2>$O/sweep_train_n2.err | tee $O/sweep_train_n2.csv | tail -20
tail -3 $O/sweep_train_n2.err
