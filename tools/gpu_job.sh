set -x
mkdir -p gpurun_out/r2w
cd $GRAFT_REPO_ROOT
O=gpurun_out/r2w
T=/tmp/ncu_r2w
mkdir -p $T
timeout 600 python bench.py --no-other-workloads --steps 2 --warmup 3 --e2e-steps 1 --no-cpu-baseline > $O/bench_short.log 2>&1 || exit 1
for wl in single genome end; do
  timeout 900 ncu --set full --clock-control none -k regex:dpc_solve_kernel -c 40 -f -o $T/full_$wl python bench.py --kernel-only --workload $wl --steps 1 --warmup 3 > $O/ncu_f_$wl.log 2>&1
  ncu -i $T/full_$wl.ncu-rep --page raw --csv > $O/raw_$wl.csv 2>$O/raw_$wl.err
done
# source pages of the two main instantiations (narrow single-gap kernel, narrow genome-gap kernel): one launch of the last pass each
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:ILi0ELi0ELb0ELi1E -s 6 -c 1 -f -o $T/src_single python bench.py --kernel-only --workload single --steps 1 --warmup 3 > $O/ncu_s_single.log 2>&1
ncu -i $T/src_single.ncu-rep --page source --csv > $O/src_single.csv 2>$O/src_single.err
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:ILi0ELi1ELb0ELi1E -s 3 -c 1 -f -o $T/src_genome python bench.py --kernel-only --workload genome --steps 1 --warmup 3 > $O/ncu_s_genome.log 2>&1
ncu -i $T/src_genome.ncu-rep --page source --csv > $O/src_genome.csv 2>$O/src_genome.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $O/launches_r2_bench.csv python bench.py --no-other-workloads --steps 2 --warmup 3 --e2e-steps 1 --no-cpu-baseline > $O/ncu_bench.log 2>&1
ls -la $O $T
du -sh $O
