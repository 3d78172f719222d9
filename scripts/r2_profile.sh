# round-2 profile pass (one GPU): ncu --set full of the top kernels + the launch list of the default bench command
set -x
B="python bench.py --no-secondary --no-e2e --no-cpu --sustain-s 0"
$B --workload seeds --groups 2000000 --chains 128 --steps 2 --warmup 1 > gpurun_out/r2_prof_seeds_plain.json 2>/dev/null || exit 1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:plate_kernel -c 1 -f -o gpurun_out/r2_seeds_v6 $B --workload seeds --groups 2000000 --chains 128 --steps 2 --warmup 1 > gpurun_out/r2_seeds_v6_ncu.log 2>&1
$B --workload rats --chains 1024 --steps 2 --warmup 1 > gpurun_out/r2_prof_rats_plain.json 2>/dev/null || exit 1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:plate_kernel -c 1 -f -o gpurun_out/r2_rats_v5 $B --workload rats --chains 1024 --steps 2 --warmup 1 > gpurun_out/r2_rats_v5_ncu.log 2>&1
python bench.py --steps 2 --warmup 1 --no-cpu --sustain-s 0 > gpurun_out/r2_prof_default_plain.json 2>/dev/null || exit 1
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_default_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e --no-secondary --sustain-s 0 > gpurun_out/r2_default_launches.log 2>&1
$B --workload pumps --steps 5 > /dev/null 2>&1 && timeout 600 ncu --set full --clock-control none --import-source on -k regex:fused_small -c 1 -f -o gpurun_out/r2_pumps_fused $B --workload pumps --steps 5 > gpurun_out/r2_pumps_fused_ncu.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -5
