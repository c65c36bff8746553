#!/bin/bash
# One gpurun call = several measurements (the GPU budget is charged per call overhead too).  Everything lands in gpurun_out/.
# usage: benchmarks/gpu_session.sh <tag> [steps...]   steps: info tests bench_c4 bench_c2 ref walk
tag=$1; shift
out=gpurun_out/$tag; mkdir -p $out
for s in "$@"; do
  case $s in
    info) (nproc; free -g; lscpu | grep -E "Model name|Socket|NUMA node\(s\)"; nvidia-smi --query-gpu=name,memory.total --format=csv) > $out/info.txt 2>&1 ;;
    tests) (time timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -40) > $out/tests.txt 2>&1 ;;
    tests_all) (time timeout 900 python -m pytest tests -m gpu -q --timeout 180 --timeout-method=thread 2>&1 | tail -80) > $out/tests.txt 2>&1 ;;
    bench_c4) (time timeout 900 python bench.py --steps 20 --warmup 5 > $out/bench_c4.json 2> $out/bench_c4.err) > $out/bench_c4.time 2>&1 ;;
    bench_c4_nocpu) (time timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $out/bench_c4.json 2> $out/bench_c4.err) > $out/bench_c4.time 2>&1 ;;
    bench_c2) (time timeout 600 python bench.py --config c2 --steps 64 --warmup 16 --no-cpu-baseline > $out/bench_c2.json 2> $out/bench_c2.err) > $out/bench_c2.time 2>&1 ;;
    ref) (time timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > $out/ref_c4.json 2> $out/ref_c4.err) > $out/ref.time 2>&1 ;;
    walk) (time timeout 900 python bench.py --config c3 --steps 8 --warmup 2 > $out/walk.json 2> $out/walk.err) > $out/walk.time 2>&1 ;;
    ncu_walk) (timeout 900 ncu --set full --import-source on -k regex:k_walk_prefix --launch-skip 30 -c 2 -f -o $out/walk_prefix python bench.py --config c3 --steps 1 --warmup 1 --no-cpu-baseline --no-gate --no-graphs --lanes 1 > $out/ncu_walk.json 2> $out/ncu_walk.err; ncu -i $out/walk_prefix.ncu-rep --page source --csv > $out/walk_prefix_source.csv 2>/dev/null; ncu -i $out/walk_prefix.ncu-rep --page raw --csv > $out/walk_prefix_raw.csv 2>/dev/null) ;;
    launches_c4) (timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $out/launches_c4.csv python bench.py --steps 8 --warmup 4 --no-cpu-baseline --no-e2e-host --no-gate > $out/launches_c4.json 2> $out/launches_c4.err) ;;
    ncu_c4) (timeout 1200 ncu --set full --import-source on --clock-control none -k regex:"k_sage_mean|k_feature|k_sample|k_prepare" --launch-skip 18 -c 9 -f -o $out/c4_top python bench.py --steps 3 --warmup 3 --lanes 1 --no-graphs --no-cpu-baseline --no-e2e-host --no-gate > $out/ncu_c4.json 2> $out/ncu_c4.err; ncu -i $out/c4_top.ncu-rep --page raw --csv > $out/c4_top_raw.csv 2>/dev/null) ;;
    sharded2) (time EU_SYM_TIMEOUT_S=10 timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tests/run_sharded_gpu.py > $out/sharded2.txt 2>&1) > $out/sharded2.time 2>&1 ;;
    bench_n2) (time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 2 --steps 20 --warmup 5 > $out/bench_n2.json 2> $out/bench_n2.err) > $out/bench_n2.time 2>&1 ;;
    bench_n2_sharded) (time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29545 bench.py --gpus 2 --steps 20 --warmup 5 --features sharded > $out/bench_n2_sharded.json 2> $out/bench_n2_sharded.err) > $out/bench_n2_sharded.time 2>&1 ;;
    bench_g*) N=${s#bench_g}; (time timeout 700 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29546 bench.py --gpus $N --steps 20 --warmup 5 > $out/bench_g$N.json 2> $out/bench_g$N.err) > $out/bench_g$N.time 2>&1 ;;
    ref_g*) N=${s#ref_g}; (time timeout 700 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29547 bench.py --impl reference --gpus $N --steps 20 --warmup 5 > $out/ref_g$N.json 2> $out/ref_g$N.err) > $out/ref_g$N.time 2>&1 ;;
    lanes_g*) N=${s#lanes_g}
        for LN in 10 20; do
          (timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29548 bench.py --gpus $N --steps 20 --warmup 5 --lanes $LN --no-gate > $out/lanes_g${N}_$LN.json 2> $out/lanes_g${N}_$LN.err)
        done ;;
    lanes_n1) for LN in 10 20; do
          (timeout 300 python bench.py --steps 20 --warmup 5 --lanes $LN --no-cpu-baseline --no-e2e-host --no-gate > $out/lanes_n1_$LN.json 2> $out/lanes_n1_$LN.err)
        done ;;
    sharded1) (EU_BENCH_FORCE_SHARDED=1 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 1 --master-addr 127.0.0.1 --master-port 29549 bench.py --gpus 1 --steps 20 --warmup 5 --no-gate > $out/sharded1.json 2> $out/sharded1.err) ;;
    launches_sharded1) (EU_BENCH_FORCE_SHARDED=1 MASTER_ADDR=127.0.0.1 MASTER_PORT=29550 RANK=0 WORLD_SIZE=1 LOCAL_RANK=0 timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum --clock-control none --launch-skip 150 -c 260 --csv --log-file $out/launches_sharded1.csv python bench.py --gpus 1 --steps 6 --warmup 3 --lanes 1 --no-graphs --no-gate > $out/launches_sharded1.json 2> $out/launches_sharded1.err) ;;
    g2_ab) run2() { tag2=$1; shift; (env "$@" timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 2 --steps 20 --warmup 5 $EXTRA > $out/g2_$tag2.json 2> $out/g2_$tag2.err); }
        EXTRA="" run2 default X=1
        EXTRA="--no-gate" run2 symctas4 EU_SYM_CTAS=4
        EXTRA="--no-gate" run2 sample6 EU_SAMPLE_CTAS=6
        EXTRA="--no-gate" run2 sym2_sample6 EU_SYM_CTAS=2 EU_SAMPLE_CTAS=6
        EXTRA="--no-gate --features sharded" run2 shardedfeat X=1 ;;
    smoke) (time timeout 300 python -c 'import __graft_entry__ as g; g.smoke()' > $out/smoke.txt 2>&1) > $out/smoke.time 2>&1 ;;
    g2_noise) run2() { tag2=$1; shift; (env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29552 bench.py --gpus 2 --steps 20 --warmup 5 --no-gate > $out/g2n_$tag2.json 2> $out/g2n_$tag2.err); }
        run2 default_a X=1; run2 knob_a EU_SYM_CTAS=2 EU_SAMPLE_CTAS=6; run2 default_b X=1; run2 knob_b EU_SYM_CTAS=2 EU_SAMPLE_CTAS=6 ;;
    ab) Q="--steps 40 --warmup 8 --no-cpu-baseline --no-e2e-host --no-gate"
        run() { tag2=$1; shift; (env "$@" timeout 300 python bench.py $Q $EXTRA > $out/ab_$tag2.json 2> $out/ab_$tag2.err); }
        EXTRA="" run c4_default X=1
        EXTRA="" run c4_nostage EU_SAMPLE_STAGE=0
        EXTRA="--lanes 8" run c4_lanes8 X=1
        EXTRA="" run c4_part_a EU_SAGE_CTAS=3 EU_SAMPLE_CTAS=6
        EXTRA="" run c4_part_b EU_SAGE_CTAS=2 EU_FEATURE_CTAS=2 EU_SAMPLE_CTAS=6
        EXTRA="--lanes 8" run c4_part_c EU_SAGE_CTAS=2 EU_FEATURE_CTAS=2 EU_SAMPLE_CTAS=6
        EXTRA="--config c2 --steps 128 --warmup 32" run c2_default X=1
        EXTRA="--config c2 --steps 128 --warmup 32" run c2_nostage EU_SAMPLE_STAGE=0 ;;
    walk_ab) (timeout 300 python bench.py --config c3 --steps 16 --warmup 4 --no-cpu-baseline > $out/walk_l8.json 2> $out/walk_l8.err)
        (timeout 300 python bench.py --config c3 --steps 32 --warmup 4 --no-cpu-baseline --no-gate --lanes 16 > $out/walk_l16.json 2> $out/walk_l16.err)
        (timeout 300 python bench.py --config c3 --steps 16 --warmup 4 --no-cpu-baseline --no-gate --rng philox > $out/walk_philox.json 2> $out/walk_philox.err) ;;
    tune) Q="--steps 40 --warmup 8 --no-cpu-baseline --no-e2e-host --no-gate --lanes 8"
        for cfg in "3 2 6" "3 3 5" "2 2 5" "4 2 4" "3 2 5" "2 3 6" "2 2 8"; do
          set -- $cfg
          (EU_SAGE_CTAS=$1 EU_FEATURE_CTAS=$2 EU_SAMPLE_CTAS=$3 timeout 300 python bench.py $Q > $out/tune_$1_$2_$3.json 2> $out/tune_$1_$2_$3.err)
        done ;;
    c5) (time timeout 600 python bench.py --config c5 --steps 32 --warmup 8 > $out/c5.json 2> $out/c5.err) > $out/c5.time 2>&1 ;;
    *) echo "unknown step $s" ;;
  esac
  echo "== $s done rc=$?" >> $out/steps.log
done
tail -c 400 $out/*.err 2>/dev/null | tail -30
cat $out/tests.txt 2>/dev/null | tail -15
