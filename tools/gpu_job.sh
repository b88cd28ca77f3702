set -x
mkdir -p gpurun_out/r2x
cd $GRAFT_REPO_ROOT
O=gpurun_out/r2x
for wl in single genome end; do
  timeout 300 python bench.py --kernel-only --workload $wl --steps 5 --warmup 3 > $O/kernel_only_$wl.log 2>&1
done
DPC_TIMING=1 timeout 300 python bench.py --no-other-workloads --steps 2 --warmup 3 --e2e-steps 4 --no-cpu-baseline > $O/bench_short.log 2> $O/bench_short.err
timeout 1800 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -4 $O/pytest.log
