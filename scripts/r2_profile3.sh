# round-2 third profile pass (one GPU): ncu --set full of the dense-predictor kernels of the final tree (reverse GEMM with
# 32-deep slabs; round 1 captured the 16-deep ones) and of the Seeds plate kernel at the bench's chain count
set -x
B="python bench.py --no-secondary --no-e2e --no-cpu --sustain-s 0"
$B --workload dense --groups 1048576 --chains 512 --steps 2 --warmup 1 > gpurun_out/r2_prof_dense_plain.json 2>gpurun_out/r2_prof_dense_plain.err || exit 1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:dense_bwd_kernel --launch-skip 1 -c 1 -f -o gpurun_out/r2_dense_bwd $B --workload dense --groups 1048576 --chains 512 --steps 2 --warmup 1 > gpurun_out/r2_dense_bwd_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:dense_fwd_kernel --launch-skip 1 -c 1 -f -o gpurun_out/r2_dense_fwd $B --workload dense --groups 1048576 --chains 512 --steps 2 --warmup 1 > gpurun_out/r2_dense_fwd_ncu.log 2>&1
ls -la gpurun_out/r2_dense_*.ncu-rep
tail -2 gpurun_out/r2_prof_dense_plain.json | cut -c1-600
