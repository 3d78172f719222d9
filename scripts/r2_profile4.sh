# round-2 last profile pass (one GPU, final tree): the ncu launch list of the default bench command and ncu --set full of the
# Rats plate kernel (1024 chains) and of the Seeds plate kernel (128 chains), each after the same command ran clean without ncu
set -x
B="python bench.py --no-secondary --no-e2e --no-cpu --sustain-s 0"
$B --steps 2 --warmup 1 > gpurun_out/r2_final_prof_plain.json 2>/dev/null || exit 1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_final_launches.csv $B --steps 2 --warmup 1 > gpurun_out/r2_final_launches.log 2>&1
$B --workload rats --chains 1024 --steps 2 --warmup 1 > /dev/null 2>&1 || exit 1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:plate_kernel -c 1 -f -o gpurun_out/r2_final_rats $B --workload rats --chains 1024 --steps 2 --warmup 1 > gpurun_out/r2_final_rats_ncu.log 2>&1
$B --workload seeds --groups 2000000 --chains 128 --steps 2 --warmup 1 > /dev/null 2>&1 || exit 1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:plate_kernel -c 1 -f -o gpurun_out/r2_final_seeds $B --workload seeds --groups 2000000 --chains 128 --steps 2 --warmup 1 > gpurun_out/r2_final_seeds_ncu.log 2>&1
ls -la gpurun_out/r2_final_*.ncu-rep gpurun_out/r2_final_launches.csv
