# Round profile set (run under gpurun from the repo root): bench lines of every configuration, the reference arm, the ncu
# launch list of a plain bench run and a full ncu capture of ONE timed step of the same command.  Output: gpurun_out/$R/
export R=${R:-r5}
set -x
mkdir -p gpurun_out/$R
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/$R/smoke.log 2>&1; tail -2 gpurun_out/$R/smoke.log
python bench.py --steps 20 --warmup 3 > gpurun_out/$R/bench_n1.json 2> gpurun_out/$R/bench_n1.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/$R/bench_reference_n1.json 2> gpurun_out/$R/bench_reference_n1.err
for c in cfg1 cfg3 cfg4 cfg5; do python bench.py --config $c --steps 3 --warmup 3 --no-cpu > gpurun_out/$R/bench_$c.json 2> gpurun_out/$R/bench_$c.err; done
python bench.py --steps 20 --warmup 3 --no-extra --no-cpu --lanes 2 > gpurun_out/$R/bench_n1_lanes2.json 2> gpurun_out/$R/bench_n1_lanes2.err
python bench.py --steps 2 --warmup 3 --no-extra --no-cpu > gpurun_out/$R/plain.json 2> gpurun_out/$R/plain.err && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/$R/launches.csv python bench.py --steps 2 --warmup 3 --no-extra --no-cpu > gpurun_out/$R/ncu_launches.log 2>&1
# one device-resident step = the launches from a clear_tables_kernel to the next fuse_kernel with 256 frames; take the fifth
# (three warm-up steps and the first timed one precede it)
SKIP=$(python - <<'P'
import csv, os
rows = list(csv.reader(open(f"gpurun_out/{os.environ.get('R', 'r5')}/launches.csv")))
h = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
rows = rows[h + 1:]
starts = [k for k, r in enumerate(rows) if "clear_tables_kernel" in r[4] and k + 1 < len(rows) and "leaf_band_kernel" in rows[k + 1][4] and rows[k + 1][8].endswith(", 256, 1)")]
k = starts[4] if len(starts) > 4 else starts[-1]
end = next(j for j in range(k, len(rows)) if "fuse_kernel" in rows[j][4])
print(k, end - k + 1)
P
)
set -- $SKIP
timeout 900 ncu --set full --clock-control none --import-source on --launch-skip $1 --launch-count $2 -o gpurun_out/$R/step_full -f python bench.py --steps 2 --warmup 3 --no-extra --no-cpu > gpurun_out/$R/ncu_full.log 2>&1
tail -3 gpurun_out/$R/ncu_full.log
