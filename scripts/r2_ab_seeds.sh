rm -f gpurun_out/r2_seeds_ab2.txt
for v in u4_mb4_e0 u2_mb8_pf u2_mb6_pf u4_mb4_pf u4_mb5_pf u2_mb8_l2pf16 u2_mb8_l2pf64 u4_mb4_l2pf32; do
  JBUGS_CUDA_LIB=$PWD/build/ab/libjbugs_$v.so python bench.py --workload seeds --no-secondary --no-e2e --no-cpu --steps 30 --warmup 3 --sustain-s 0 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$v', round(d['ms_per_step'],3), round(d['roofline']['kernel_ms'],3), d['clocks']['sm_mhz'])" >> gpurun_out/r2_seeds_ab2.txt
done
cat gpurun_out/r2_seeds_ab2.txt
python -m pytest tests -m gpu -x -q > gpurun_out/r2_gputest_b.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_gputest_b.log; tail -5 gpurun_out/r2_gputest_b.log
