set -x
mkdir -p gpurun_out/r2u
cd $GRAFT_REPO_ROOT
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/r2u/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2u/pytest.log
tail -6 gpurun_out/r2u/pytest.log
