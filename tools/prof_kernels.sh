#!/bin/bash
# ncu evidence for profiles/: launch list of one step + full captures of the dominant kernels (run under gpurun)
TAG=${1:-r02}
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/${TAG}_launches.csv python tools/profile_step.py --what step > gpurun_out/prof_step.log 2>&1
for w in head_conv frustum_conv redir1x1 depth_conv bri enc_conv; do
  ncu --set full --clock-control none --import-source on --profile-from-start off -f -o gpurun_out/${TAG}_$w python tools/profile_step.py --what $w > gpurun_out/prof_$w.log 2>&1
  ncu -i gpurun_out/${TAG}_$w.ncu-rep --page details --csv > gpurun_out/${TAG}_${w}_details.csv 2>/dev/null
  ncu -i gpurun_out/${TAG}_$w.ncu-rep --page raw --csv > gpurun_out/${TAG}_${w}_raw.csv 2>/dev/null
  rm -f gpurun_out/${TAG}_$w.ncu-rep
done
python tools/ncu_traffic.py gpurun_out/${TAG}_head_conv_raw.csv gpurun_out/${TAG}_head_conv_traffic.json
ls gpurun_out | grep ${TAG} | head -30
