set -x
mkdir -p gpurun_out/r2r
cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2r/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2r/pytest.log
tail -4 gpurun_out/r2r/pytest.log
for lib in build/libdpc_prev.so gmap-gsnap_b200/csrc/libdynprog_cuda.so; do
  for wl in single end genome; do
  n=1000000; [ $wl = genome ] && n=500000
  echo "== $lib $wl" >> gpurun_out/r2r/kernel_only.log
  DPC_LIB=$PWD/$lib timeout 300 python bench.py --kernel-only --workload $wl --problems $n --steps 5 --warmup 3 >> gpurun_out/r2r/kernel_only.log 2>&1
  done
done
cut -c1-120 gpurun_out/r2r/kernel_only.log
