set -x
mkdir -p gpurun_out/r3h
cd $GRAFT_REPO_ROOT
O=gpurun_out/r3h
timeout 900 python bench.py > $O/bench.log 2> $O/bench.err; echo "bench rc=$?"
