#!/bin/bash
# ncu evidence for profiles/: launch list of one step + full captures of the dominant kernels (run under gpurun)
TAG=${1:-r02}
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/${TAG}_launches.csv python tools/profile_step.py --what step > gpurun_out/prof_step.log 2>&1
for spec in head_conv:f16 head_conv:tf32 enc_conv:f16 enc_conv:tf32 frustum_conv:tf32 depth_conv:tf32x3 redir1x1:tf32 bri:tf32; do
  w=${spec%%:*}; m=${spec##*:}
  ncu --set full --clock-control none --import-source on --profile-from-start off -f -o gpurun_out/${TAG}_${w}_$m python tools/profile_step.py --what $w --kernel-math $m > gpurun_out/prof_${w}_$m.log 2>&1
  ncu -i gpurun_out/${TAG}_${w}_$m.ncu-rep --page details --csv > gpurun_out/${TAG}_${w}_${m}_details.csv 2>/dev/null
  ncu -i gpurun_out/${TAG}_${w}_$m.ncu-rep --page raw --csv > gpurun_out/${TAG}_${w}_${m}_raw.csv 2>/dev/null
  rm -f gpurun_out/${TAG}_${w}_$m.ncu-rep
done
python tools/ncu_traffic.py gpurun_out/${TAG}_head_conv_f16_raw.csv gpurun_out/${TAG}_head_conv_f16_traffic.json
python tools/ncu_traffic.py gpurun_out/${TAG}_head_conv_tf32_raw.csv gpurun_out/${TAG}_head_conv_tf32_traffic.json
ls gpurun_out | grep ${TAG} | head -40
