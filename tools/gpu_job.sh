set -x
mkdir -p gpurun_out/r3f
cd $GRAFT_REPO_ROOT
O=gpurun_out/r3f
for wl in genome single; do
  timeout 300 python bench.py --kernel-only --workload $wl --steps 5 --warmup 3 > $O/kernel_only_$wl.log 2>&1
done
timeout 1800 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest.log
