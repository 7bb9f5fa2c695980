#!/bin/bash
# round 2, GPU call 32 (2 GPUs): sanity of the multi-rank bench path on the final tree
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
T=gpurun_out/r2_32
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29602 bench.py --gpus 2 --steps 5 --warmup 3 > ${T}_bench_c2_2gpu.json 2> ${T}_bench_c2_2gpu.err; echo "bench c2 N=2 exit $?"; cut -c1-220 ${T}_bench_c2_2gpu.json
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29603 bench.py --gpus 2 --impl reference --steps 1 --warmup 1 > ${T}_bench_ref_2gpu.json 2> ${T}_bench_ref_2gpu.err; echo "bench reference N=2 exit $?"; cut -c1-160 ${T}_bench_ref_2gpu.json
timeout 300 python -m pytest -m gpu -q --timeout 200 tests/test_nccl_combine.py tests/test_gpu_parity.py -k "multi_gpu or nccl or allreduce" > ${T}_tests.log 2>&1; echo "pytest exit $?"; tail -2 ${T}_tests.log
