#!/bin/bash
# bench lines for the main workloads + ncu launch list + one full capture.   gpurun --timeout 1200 -- bash tools/gpu_bench.sh <tag> [ncu]
TAG=${1:-x}
mkdir -p gpurun_out
python bench.py --no-cpu-baseline > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"; cat gpurun_out/bench_$TAG.json; tail -5 gpurun_out/bench_$TAG.err
for wl in cfg4main cfg5 cfg2 cfg3; do python bench.py --workload $wl --no-cpu-baseline > gpurun_out/bench_${wl}_$TAG.json 2>> gpurun_out/bench_$TAG.err; python - <<PY
import json
d=json.load(open("gpurun_out/bench_${wl}_$TAG.json"))
print("$wl", d["value"], "img/s", d["ms_per_step"], "ms/step", d["roofline"]["kernel_ms"], "step_frac", d["roofline"]["step_frac"], "e2e", d["e2e"]["value"])
PY
done
if [ "$2" == "ncu" ]; then
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/plain_$TAG.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_launches_$TAG.log 2>&1; echo "ncu launches rc=$?"
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/plain2_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:decode_planes -s 4 -c 1 -o gpurun_out/prof_$TAG -f \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_full_$TAG.log 2>&1; echo "ncu full rc=$?"
fi
