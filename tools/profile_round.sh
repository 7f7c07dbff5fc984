set -x
mkdir -p gpurun_out/r4
python bench.py --steps 20 --warmup 3 > gpurun_out/r4/bench_n1.json 2> gpurun_out/r4/bench_n1.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r4/bench_reference_n1.json 2> gpurun_out/r4/bench_reference_n1.err
for c in cfg1 cfg3 cfg4 cfg5; do python bench.py --config $c --steps 3 --warmup 3 --no-cpu > gpurun_out/r4/bench_$c.json 2> gpurun_out/r4/bench_$c.err; done
python bench.py --steps 2 --warmup 3 --no-extra --no-cpu > gpurun_out/r4/plain.json 2> gpurun_out/r4/plain.err && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r4/launches.csv python bench.py --steps 2 --warmup 3 --no-extra --no-cpu > gpurun_out/r4/ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on --launch-skip 71 --launch-count 23 -o gpurun_out/r4/step_full -f python bench.py --steps 2 --warmup 3 --no-extra --no-cpu > gpurun_out/r4/ncu_full.log 2>&1
tail -3 gpurun_out/r4/ncu_full.log
