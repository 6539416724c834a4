#!/bin/bash
# One call on an 8-GPU box: how the host feeds 8 ranks (feed probe), the NCCL rank tests on all GPUs, bench at N = 8 and 4.
# Every step has its own timeout; outputs under gpurun_out/ with the prefix given as $1.
P=${1:-r2k}
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
nproc; free -g | head -2
timeout 300 $TR --nproc-per-node 8 --master-port 29521 profiles/feed_probe.py 100000 > gpurun_out/${P}_feed_n8.txt 2>&1; echo "feed rc=$?"
grep "PP_FEED" gpurun_out/${P}_feed_n8.txt | sort | awk '{print $1,$2,$3,$4,$5}' | head -80
timeout 400 $TR --nproc-per-node 8 --master-port 29522 bench.py --gpus 8 > gpurun_out/${P}_bench_n8.json 2> gpurun_out/${P}_bench_n8.err; echo "bench8 rc=$?"
PP_FEED=zerocopy timeout 200 $TR --nproc-per-node 8 --master-port 29523 bench.py --gpus 8 --workloads none --quick > gpurun_out/${P}_bench_n8_zerocopy.json 2> gpurun_out/${P}_bench_n8_zerocopy.err; echo "bench8 zc rc=$?"
timeout 300 python -m pytest tests/test_gpu_dist.py -m gpu -x -q > gpurun_out/${P}_pytest_dist.txt 2>&1; echo "dist tests rc=$?"; tail -n 3 gpurun_out/${P}_pytest_dist.txt
timeout 300 $TR --nproc-per-node 4 --master-port 29524 bench.py --gpus 4 > gpurun_out/${P}_bench_n4.json 2> gpurun_out/${P}_bench_n4.err; echo "bench4 rc=$?"
