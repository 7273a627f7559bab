M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,lts__t_bytes.sum,sm__cycles_elapsed.avg.per_second
python scripts/dev_rd_once.py > gpurun_out/prof_rd_plain.log 2>&1 || exit 1
ncu --profile-from-start off --metrics $M --clock-control none --csv --log-file gpurun_out/r02g_launches_rdresunet.csv python scripts/dev_rd_once.py > gpurun_out/prof_ncu6.log 2>&1
bash scripts/prof_rd_aux.sh r02g > gpurun_out/prof_rd_aux.log 2>&1
ls gpurun_out | grep r02g
