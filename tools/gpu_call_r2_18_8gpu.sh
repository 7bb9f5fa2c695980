#!/bin/bash
# round 2, GPU call 18 (8 GPUs): final tree -- weak scaling of C2 and the C3 target run at N = 8, 4, 2; multi-GPU tests
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
T=gpurun_out/r2_18
run() { # N workload steps warmup
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $((29500 + $1)) bench.py --gpus $1 --workload $2 --steps $3 --warmup $4 > ${T}_bench_$2_$1gpu.json 2> ${T}_bench_$2_$1gpu.err; echo "bench $2 N=$1 exit $?"; cut -c1-260 ${T}_bench_$2_$1gpu.json
}
run 8 c2 5 3
run 8 c3 3 3
run 4 c2 5 3
run 4 c3 3 3
run 2 c2 5 3
run 2 c3 2 3
timeout 300 python -m pytest -m gpu -v --timeout 200 tests/test_nccl_combine.py tests/test_gpu_parity.py -k "multi_gpu or nccl or allreduce" > ${T}_multi_gpu_tests.log 2>&1; echo "pytest exit $?"; tail -6 ${T}_multi_gpu_tests.log
