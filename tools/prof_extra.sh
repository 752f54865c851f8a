for w in stage2_conv tpose32 hg_conv1 neck_k4 hg_redir2; do
  ncu --set full --clock-control none --profile-from-start off -f -o gpurun_out/x_$w python tools/profile_step.py --what $w > gpurun_out/prof_$w.log 2>&1
  ncu -i gpurun_out/x_$w.ncu-rep --page details --csv > gpurun_out/x_${w}_details.csv 2>/dev/null
  ncu -i gpurun_out/x_$w.ncu-rep --page raw --csv > gpurun_out/x_${w}_raw.csv 2>/dev/null
  rm -f gpurun_out/x_$w.ncu-rep
done
