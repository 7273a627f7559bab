# Profiling pass of one bench step (run under gpurun): plain run first (must exit 0), then the launch list with DRAM bytes +
# tensor-pipe activity, then full captures of the dominant launches.  bench.py --kernels-only brackets the timed steps with
# cudaProfilerStart/Stop, so --profile-from-start off captures exactly those launches.   usage: scripts/prof_step.sh r02
R=${1:-r02}
set -x
CMD="python bench.py --steps 2 --warmup 3 --kernels-only --no-sub"
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,lts__t_bytes.sum,sm__cycles_elapsed.avg.per_second
$CMD > gpurun_out/prof_plain.log 2>&1 || exit 1
ncu --profile-from-start off --metrics $M --clock-control none --csv --log-file gpurun_out/${R}_launches_step.csv $CMD > gpurun_out/prof_ncu1.log 2>&1
cat gpurun_out/prof_plain.log
tail -n 3 gpurun_out/prof_ncu1.log
# full captures (one launch each; ncu replays the kernel ~40 times)
ncu --profile-from-start off --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:conv_v3_kernelILi1ELi1ELb0ELb1E -c 1 -f -o gpurun_out/${R}_v3_recon_pre $CMD > gpurun_out/prof_ncu2.log 2>&1
ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:metric_kernel -c 1 -f -o gpurun_out/${R}_metric $CMD > gpurun_out/prof_ncu3.log 2>&1
ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:crappify_kernel -c 1 -f -o gpurun_out/${R}_crappify $CMD > gpurun_out/prof_ncu4.log 2>&1
python scripts/dev_stitch_once.py > gpurun_out/prof_stitch_plain.log 2>&1 &&
ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:stitch_band16 -c 1 -f -o gpurun_out/${R}_stitch python scripts/dev_stitch_once.py > gpurun_out/prof_ncu5.log 2>&1
# RDResUNet forward (batch 50): launch list
python scripts/dev_rd_once.py > gpurun_out/prof_rd_plain.log 2>&1 &&
ncu --profile-from-start off --metrics $M --clock-control none --csv --log-file gpurun_out/${R}_launches_rdresunet.csv python scripts/dev_rd_once.py > gpurun_out/prof_ncu6.log 2>&1
ls -la gpurun_out/ | tail -20
