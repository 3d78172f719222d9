rm -f gpurun_out/r2_seeds_ab4.txt
for v in 2_8_4_1 2_8_8_1 2_8_16_1 4_4_4_1 4_4_8_1 4_4_16_1 4_4_8_0; do
  JBUGS_CUDA_LIB=$PWD/build/ab/libjbugs_$v.so python bench.py --workload seeds --no-secondary --no-e2e --no-cpu --steps 30 --warmup 3 --sustain-s 0 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$v', round(d['ms_per_step'],3), round(d['roofline']['kernel_ms'],3), d['clocks']['sm_mhz'])" >> gpurun_out/r2_seeds_ab4.txt
done
cat gpurun_out/r2_seeds_ab4.txt
python -m pytest tests/test_gpu_edge.py tests/test_abi.py -m gpu -x -q > gpurun_out/r2_gputest_d.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_gputest_d.log; tail -5 gpurun_out/r2_gputest_d.log
