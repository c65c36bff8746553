#!/bin/bash
# One gpurun call = several measurements (the GPU budget is charged per call overhead too).  Everything lands in gpurun_out/.
# usage: benchmarks/gpu_session.sh <tag> [steps...]   steps: info tests bench_c4 bench_c2 ref walk
tag=$1; shift
out=gpurun_out/$tag; mkdir -p $out
for s in "$@"; do
  case $s in
    info) (nproc; free -g; lscpu | grep -E "Model name|Socket|NUMA node\(s\)"; nvidia-smi --query-gpu=name,memory.total --format=csv) > $out/info.txt 2>&1 ;;
    tests) (time timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -40) > $out/tests.txt 2>&1 ;;
    tests_all) (time timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -60) > $out/tests.txt 2>&1 ;;
    bench_c4) (time timeout 900 python bench.py --steps 20 --warmup 5 > $out/bench_c4.json 2> $out/bench_c4.err) > $out/bench_c4.time 2>&1 ;;
    bench_c4_nocpu) (time timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $out/bench_c4.json 2> $out/bench_c4.err) > $out/bench_c4.time 2>&1 ;;
    bench_c2) (time timeout 600 python bench.py --config c2 --steps 64 --warmup 16 --no-cpu-baseline > $out/bench_c2.json 2> $out/bench_c2.err) > $out/bench_c2.time 2>&1 ;;
    ref) (time timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > $out/ref_c4.json 2> $out/ref_c4.err) > $out/ref.time 2>&1 ;;
    walk) (time timeout 900 python bench.py --config c3 --steps 8 --warmup 2 > $out/walk.json 2> $out/walk.err) > $out/walk.time 2>&1 ;;
    *) echo "unknown step $s" ;;
  esac
  echo "== $s done rc=$?" >> $out/steps.log
done
tail -c 400 $out/*.err 2>/dev/null | tail -30
cat $out/tests.txt 2>/dev/null | tail -15
