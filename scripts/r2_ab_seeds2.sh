# build Seeds kernel variants (U_MINB_PF_L2PF) into build/ab/libjbugs_<v>.so; run on the GPU box with: bash scripts/r2_ab_seeds2.sh run
cd "$(dirname "$0")/../juliabugs.jl_b200/csrc"
VARS="4_3_true_8 3_4_true_8 6_3_false_8"
if [ "$1" = "build" ]; then
  mkdir -p ../../build/ab
  for v in $VARS; do
    IFS=_ read u mb pf l2 <<< "$v"
    ( nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo --expt-relaxed-constexpr -Xcompiler -fPIC -ccbin g++ -Xptxas -v \
        '-DJB_SPEC=StaticSpec<SeedsDesc>' -DJB_BLK=128 -DJB_TMA=false -DJB_NAME=launch_seeds_128 -DJB_SEEDS_UNROLL=$u -DJB_SEEDS_MINB=$mb -DJB_SEEDS_PF=$pf -DJB_SEEDS_L2PF=$l2 \
        -c -o ../../build/ab/seeds_$v.o spec_one.cu 2> ../../build/ab/seeds_$v.log
      objs=$(ls *.o | grep -v '^spec_seeds_128.o$' | tr '\n' ' ')
      nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../../build/ab/libjbugs_$v.so $objs ../../build/ab/seeds_$v.o -ldl -lpthread
      echo "$v $(grep -o 'Used [0-9]* registers' ../../build/ab/seeds_$v.log | head -1) $(grep -o '[0-9]* bytes spill stores' ../../build/ab/seeds_$v.log | head -1)" ) &
  done
  wait
else
  cd ../..
  rm -f gpurun_out/r2_seeds_ab5.txt
  for v in $VARS; do
    JBUGS_CUDA_LIB=$PWD/build/ab/libjbugs_$v.so python bench.py --workload seeds --no-secondary --no-e2e --no-cpu --steps 30 --warmup 3 --sustain-s 0 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$v', round(d['ms_per_step'],3), round(d['roofline']['kernel_ms'],3), d['clocks']['sm_mhz'])" >> gpurun_out/r2_seeds_ab5.txt
  done
  cat gpurun_out/r2_seeds_ab5.txt
fi
