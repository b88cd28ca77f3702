set -x
mkdir -p gpurun_out/r3e
cd $GRAFT_REPO_ROOT
O=gpurun_out/r3e
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $O/smoke.log
timeout 1800 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest.log
timeout 900 python bench.py --workload gmap > $O/bench_gmap.log 2> $O/bench_gmap.err; echo "gmap rc=$?"; tail -c 1500 $O/bench_gmap.log
