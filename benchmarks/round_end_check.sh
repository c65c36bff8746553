set -x
python -m pytest tests -x -q -m gpu 2>&1 | tail -3
python bench.py > gpurun_out/bench_n1_default.json 2> gpurun_out/bench_n1_default.err; tail -c 600 gpurun_out/bench_n1_default.err
python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/bench_n1_k50.json 2> gpurun_out/bench_n1_k50.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; tail -c 300 gpurun_out/bench_ref.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r01_launches_final.csv python bench.py --steps 32 --warmup 16 --lanes 1 --no-cpu-baseline > gpurun_out/ncu_b.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:'k_sample|k_sage_mean|k_feature|k_prepare' -s 12 -c 8 -f -o gpurun_out/prof_r1_final python bench.py --steps 32 --warmup 16 --lanes 1 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out | tail -8
