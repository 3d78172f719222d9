# round-2 second profile pass: the HMC probe on the final tree, ncu --set full of the run-time compiled Rats-variant kernel (best variant),
# of a compiled tape kernel and of the generated-quantities kernel
set -x
python scripts/hmc_probe.py > gpurun_out/r2_hmc_probe.txt 2>&1
B="python bench.py --no-secondary --no-e2e --no-cpu --sustain-s 0"
$B --workload rats_variant --steps 3 > /dev/null 2>&1 || exit 1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:plate_kernel --launch-skip 8 -c 1 -f -o gpurun_out/r2_jit_rats_best $B --workload rats_variant --steps 3 > gpurun_out/r2_jit_rats_best_ncu.log 2>&1
$B --workload mixture --steps 3 > /dev/null 2>&1 || exit 1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:plate_kernel --launch-skip 3 -c 1 -f -o gpurun_out/r2_tape_mixture $B --workload mixture --steps 3 > gpurun_out/r2_tape_mixture_ncu.log 2>&1
ls -la gpurun_out/r2_jit_rats_best.ncu-rep gpurun_out/r2_tape_mixture.ncu-rep
tail -8 gpurun_out/r2_hmc_probe.txt
