# profiling pass of one bench step (run under gpurun): plain run, launch list with DRAM bytes + tensor-pipe activity, and
# full captures of the dominant conv launches.  bench.py --kernels-only brackets the timed steps with cudaProfilerStart/Stop,
# so --profile-from-start off captures exactly those launches.
set -x
CMD="python bench.py --steps 2 --warmup 3 --kernels-only"
$CMD > gpurun_out/prof_plain.log 2>&1 || exit 1
ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,lts__t_bytes.sum,sm__cycles_elapsed.avg.per_second \
    --clock-control none --csv --log-file gpurun_out/r01_launches_v3.csv $CMD > gpurun_out/prof_ncu1.log 2>&1
cat gpurun_out/prof_plain.log
tail -n 3 gpurun_out/prof_ncu1.log
