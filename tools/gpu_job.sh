set -x
mkdir -p gpurun_out/r3i
cd $GRAFT_REPO_ROOT
O=gpurun_out/r3i
T=/tmp/ncu_r3i
mkdir -p $T
timeout 600 ncu --set full --clock-control none -k regex:dpc_solve_kernel -c 40 -f -o $T/full_genome python bench.py --kernel-only --workload genome --steps 1 --warmup 3 > $O/ncu_f_genome.log 2>&1
ncu -i $T/full_genome.ncu-rep --page raw --csv > $O/raw_genome.csv 2>$O/raw_genome.err
ls -la $O
