set -x
mkdir -p gpurun_out/r2i
cd $GRAFT_REPO_ROOT
run() { name=$1; shift; env "$@" DPC_TIMING=1 timeout 600 python bench.py --no-cpu-baseline --steps 3 > gpurun_out/r2i/bench_$name.log 2> gpurun_out/r2i/bench_$name.err; echo "== $name"; grep -h "pairs 1" gpurun_out/r2i/bench_$name.err | tail -2 | cut -c1-300; grep -h "pairs 0" gpurun_out/r2i/bench_$name.err | tail -1 | cut -c1-200; }
run d1 DPC_DRIVERS=1
run d2 DPC_DRIVERS=2
run d3 DPC_DRIVERS=3
run d2s24 DPC_DRIVERS=2 DPC_PIPE_SLOTS=24
run d2c8k DPC_DRIVERS=2 DPC_CHUNK=8192
run d2c32k DPC_DRIVERS=2 DPC_CHUNK=32768
run d2host DPC_DRIVERS=2 DPC_ROUTE=host
run d2dev DPC_DRIVERS=2 DPC_ROUTE=device
run d2l1 DPC_DRIVERS=2 DPC_LINK_DEPTH=1
run d2l0 DPC_DRIVERS=2 DPC_LINK_DEPTH=0.5
