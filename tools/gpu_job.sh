set -x
mkdir -p gpurun_out/r2j
cd $GRAFT_REPO_ROOT
nproc > gpurun_out/r2j/nproc.txt
timeout 900 python bench.py > gpurun_out/r2j/bench.log 2> gpurun_out/r2j/bench.err; echo "rc=$?"
tail -5 gpurun_out/r2j/bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2j/bench_ref.log 2> gpurun_out/r2j/bench_ref.err; echo "rc=$?"
cat gpurun_out/r2j/bench_ref.log | cut -c1-600
