# ncu --set full of the RDNet encoder's CUDA-core kernels (one launch each) inside one RDResUNet forward (batch 50).
# usage (under gpurun): scripts/prof_rd_aux.sh r02
R=${1:-r02}
CMD="python scripts/dev_rd_once.py"
$CMD > gpurun_out/prof_rd_plain.log 2>&1 || exit 1
for spec in "dwconv7_kernel:12:dwconv7_small" "dwconv7_kernel:0:dwconv7_lo" "ln_kernel:3:ln_s2d" "ese_fused:0:ese_fused"; do
  IFS=: read k skip name <<< "$spec"
  ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:$k --launch-skip $skip -c 1 -f -o gpurun_out/${R}_$name $CMD > gpurun_out/prof_ncu_$name.log 2>&1
  ncu -i gpurun_out/${R}_$name.ncu-rep --page details > gpurun_out/${R}_$name.details.txt 2>&1
done
ls -la gpurun_out | tail
