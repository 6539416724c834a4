#!/bin/bash
# bench.py at N = 8 and N = 4 on one 8-GPU box (outputs under gpurun_out/ with the prefix $1)
P=${1:-r2p}
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
PP_TRACE=1 timeout 400 $TR --nproc-per-node 8 --master-port 29522 bench.py --gpus 8 > gpurun_out/${P}_bench_n8.json 2> gpurun_out/${P}_bench_n8.err; echo "bench8 rc=$?"
timeout 300 $TR --nproc-per-node 4 --master-port 29524 bench.py --gpus 4 > gpurun_out/${P}_bench_n4.json 2> gpurun_out/${P}_bench_n4.err; echo "bench4 rc=$?"
grep -c "zero-copy gather" gpurun_out/${P}_bench_n8.err
