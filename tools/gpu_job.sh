set -x
mkdir -p gpurun_out/r3d
cd $GRAFT_REPO_ROOT
O=gpurun_out/r3d
N=$(nvidia-smi -L | wc -l)
run() {
  tag=$1; shift
  env "$@" DPC_TIMING=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --no-other-workloads --no-cpu-baseline --steps 3 --e2e-steps 3 > $O/bench_n${N}_$tag.log 2> $O/bench_n${N}_$tag.err
  grep -o '"e2e": {"value": [0-9.]*, "unit": "GCUPS", "fills_per_s": [0-9.]*, "ms_per_step": [0-9.]*' $O/bench_n${N}_$tag.log
}
run hd2 DPC_HOST_DEPTH=2
run dev DPC_ROUTE=device
