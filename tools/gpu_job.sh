set -x
mkdir -p gpurun_out/r2t
cd $GRAFT_REPO_ROOT
for n in 1 2 4 8; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2954$n tools/pcie_probe.py 2>/dev/null | grep ranks >> gpurun_out/r2t/pcie.log
done
cat gpurun_out/r2t/pcie.log
DPC_TIMING=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29560 bench.py --gpus 8 --no-other-workloads --steps 3 --no-cpu-baseline > gpurun_out/r2t/bench_n8.log 2> gpurun_out/r2t/bench_n8.err
grep "dpc_solve" gpurun_out/r2t/bench_n8.err | tail -12 | cut -c1-330
for route in host device; do
DPC_ROUTE=$route timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 8 --no-other-workloads --steps 3 --no-cpu-baseline > gpurun_out/r2t/bench_n8_$route.log 2> gpurun_out/r2t/bench_n8_$route.err
python - <<PY
import json
t=open('gpurun_out/r2t/bench_n8_$route.log').read()
l=json.loads(t[t.index('{'):])
print("$route", l['e2e']['value'], l['e2e']['ms_per_step'], l['e2e']['results_only']['ms_per_step'])
PY
done
