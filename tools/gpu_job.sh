set -x
mkdir -p gpurun_out/r2d
cd $GRAFT_REPO_ROOT
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2d/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2d/pytest.log
tail -5 gpurun_out/r2d/pytest.log
for wl in single genome end; do
  n=1000000; [ $wl = genome ] && n=500000
  echo "== $wl" >> gpurun_out/r2d/kernel_only.log
  timeout 300 python bench.py --kernel-only --workload $wl --problems $n --steps 5 --warmup 3 >> gpurun_out/r2d/kernel_only.log 2>&1
done
cat gpurun_out/r2d/kernel_only.log
