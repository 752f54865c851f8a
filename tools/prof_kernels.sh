for w in redir1x1 frustum_conv depth_conv bri; do
  ncu --set full --clock-control none --import-source on --profile-from-start off -f -o gpurun_out/r01s2_$w python tools/profile_step.py --what $w > gpurun_out/prof_$w.log 2>&1
  ncu -i gpurun_out/r01s2_$w.ncu-rep --page details --csv > gpurun_out/r01s2_${w}_details.csv 2>/dev/null
done
ls -la gpurun_out | tail -12
