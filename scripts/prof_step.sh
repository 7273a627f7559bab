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
# full captures of the dominant launches (one capture each; ncu replays the kernel ~40 times)
ncu --profile-from-start off --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:conv_v3_kernelILi1ELi1ELb0ELb1E -c 1 -f -o gpurun_out/r01_v3_recon_pre $CMD > gpurun_out/prof_ncu2.log 2>&1
ncu --profile-from-start off --set full --clock-control none -k regex:metric_kernel -c 1 -f -o gpurun_out/r01_metric $CMD > gpurun_out/prof_ncu3.log 2>&1
ncu --profile-from-start off --set full --clock-control none -k regex:crappify_kernel -c 1 -f -o gpurun_out/r01_crappify $CMD > gpurun_out/prof_ncu4.log 2>&1
ncu --profile-from-start off --set full --clock-control none -k regex:tailsum -c 1 -f -o gpurun_out/r01_tailsum $CMD > gpurun_out/prof_ncu5.log 2>&1
