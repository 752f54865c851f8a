#!/bin/bash
# ncu evidence for the image encoder (row N2): launch list + full captures of selected launches (run under gpurun)
TAG=${1:-r02_image}
PROFILE_ONE=mixed ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/${TAG}_launches.csv python tools/image_encoder_timing.py > gpurun_out/prof_image.log 2>&1
# launch ids (0-based, from the launch list): override with IDS="a b c"
for id in ${IDS:-34 175 35}; do
  PROFILE_ONE=mixed ncu --set full --clock-control none --import-source on --profile-from-start off --launch-skip $id --launch-count 1 -f -o gpurun_out/${TAG}_k$id python tools/image_encoder_timing.py > gpurun_out/prof_image_k$id.log 2>&1
  ncu -i gpurun_out/${TAG}_k$id.ncu-rep --page details --csv > gpurun_out/${TAG}_k${id}_details.csv 2>/dev/null
  ncu -i gpurun_out/${TAG}_k$id.ncu-rep --page raw --csv > gpurun_out/${TAG}_k${id}_raw.csv 2>/dev/null
  rm -f gpurun_out/${TAG}_k$id.ncu-rep
done
ls gpurun_out | grep ${TAG}
