set -x
mkdir -p gpurun_out/r2o
cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -m gpu -x -q -k "genome or golden or mixed or compiled_reference" > gpurun_out/r2o/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2o/pytest.log
tail -4 gpurun_out/r2o/pytest.log
for lib in gmap-gsnap_b200/csrc/libdynprog_cuda.so build/libdpc_g3.so; do
  echo "== $lib genome" >> gpurun_out/r2o/kernel_only.log
  DPC_LIB=$PWD/$lib timeout 300 python bench.py --kernel-only --workload genome --problems 500000 --steps 5 --warmup 3 >> gpurun_out/r2o/kernel_only.log 2>&1
done
cut -c1-150 gpurun_out/r2o/kernel_only.log
