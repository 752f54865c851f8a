w=pw_proj
ncu --set full --clock-control none --import-source on --profile-from-start off -f -o gpurun_out/x_$w python tools/profile_step.py --what $w > gpurun_out/prof_$w.log 2>&1
ncu -i gpurun_out/x_$w.ncu-rep --page details --csv > gpurun_out/x_${w}_details.csv 2>/dev/null
ncu -i gpurun_out/x_$w.ncu-rep --page source --csv > gpurun_out/x_${w}_src.csv 2>/dev/null
rm -f gpurun_out/x_$w.ncu-rep
python tools/profile_step.py --what pw_proj --time-one | grep time
