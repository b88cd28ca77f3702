set -x
mkdir -p gpurun_out/r2z
cd $GRAFT_REPO_ROOT
O=gpurun_out/r2z
for mix in auto 0 35 50 65 100; do
  if [ $mix = auto ]; then unset DPC_ROUTE_MIX; else export DPC_ROUTE_MIX=$mix; fi
  DPC_TIMING=1 timeout 300 python bench.py --no-other-workloads --steps 2 --warmup 3 --e2e-steps 4 --no-cpu-baseline > $O/bench_mix_$mix.log 2> $O/bench_mix_$mix.err
  grep -o '"e2e": {"value": [0-9.]*, "unit": "GCUPS", "fills_per_s": [0-9.]*, "ms_per_step": [0-9.]*' $O/bench_mix_$mix.log
  grep "dpc_solve n=1000000" $O/bench_mix_$mix.err | tail -3 | cut -c1-330
done
