# Scratch job script for `gpurun -- 'bash tools/gpu_job.sh'` (edited per run; outputs under gpurun_out/, which is not tracked).
# Default content: the round-end sequence -- smoke, GPU parity tests, the bench line of both arms.
set -x
mkdir -p gpurun_out/job
cd $GRAFT_REPO_ROOT
O=gpurun_out/job
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $O/smoke.log
timeout 1800 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest.log
timeout 900 python bench.py --impl reference > $O/bench_ref.log 2> $O/bench_ref.err; echo "ref rc=$?"
timeout 900 python bench.py > $O/bench.log 2> $O/bench.err; echo "bench rc=$?"
