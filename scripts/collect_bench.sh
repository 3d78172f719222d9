#!/bin/bash
# One-GPU measurement sweep over the BASELINE.json configurations; JSON lines land in gpurun_out/bench_lines.jsonl.
# (run on the GPU box:  gpurun --timeout 1500 -- 'bash scripts/collect_bench.sh')
out=gpurun_out/bench_lines.jsonl
mkdir -p gpurun_out; : > $out
run() { python bench.py "$@" 2>>gpurun_out/collect.err | tail -n 1 >> $out; }
run --steps 10 --warmup 3                                        # config 2 (default line: value, roofline, cpu_baseline, e2e)
run --workload seeds --no-cpu --steps 10 --warmup 3              # config 3, 256 chains
run --workload seeds --chains 64 --no-cpu --no-e2e --steps 10 --warmup 3
run --workload pumps --no-cpu --no-e2e --steps 50 --warmup 5     # config 5
run --workload epil --no-cpu --no-e2e --steps 50 --warmup 5
run --workload dense --no-cpu --no-e2e --steps 3 --warmup 3      # config 4
python bench.py --impl reference --steps 3 --warmup 1 2>>gpurun_out/collect.err | tail -n 1 >> $out
wc -l $out
