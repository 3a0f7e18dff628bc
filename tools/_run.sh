T=r02f
set -x
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$T.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_$T.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python bench.py > gpurun_out/bench_cfg4_$T.json 2> gpurun_out/bench_cfg4_$T.err; echo "bench rc=$?"; cat gpurun_out/bench_cfg4_$T.json | head -c 600
timeout 600 python bench.py --impl reference > gpurun_out/bench_zref_$T.json 2> gpurun_out/bench_zref_$T.err; echo "ref rc=$?"; cat gpurun_out/bench_zref_$T.json | head -c 300
for wl in cfg2 cfg2x cfg3 cfg4main cfg5; do
  timeout 300 python bench.py --workload $wl --no-cpu-baseline --steps 50 > gpurun_out/bench_${wl}_$T.json 2> gpurun_out/bench_${wl}_$T.err; echo "$wl rc=$?"
done
timeout 300 python bench.py --workload cfg4 --dtype bf16 --no-cpu-baseline --steps 50 > gpurun_out/bench_cfg4bf16_$T.json 2> gpurun_out/bench_cfg4bf16_$T.err; echo "bf16 rc=$?"
timeout 300 python bench.py --workload cfg4 --graph --no-cpu-baseline --no-e2e --steps 50 > gpurun_out/bench_cfg4graph_$T.json 2> gpurun_out/bench_cfg4graph_$T.err; echo "graph rc=$?"
timeout 300 python bench.py --workload cfg2 --graph --no-cpu-baseline --no-e2e --steps 50 > gpurun_out/bench_cfg2graph_$T.json 2> gpurun_out/bench_cfg2graph_$T.err; echo "graph2 rc=$?"
timeout 300 python tools/bench_boxfit.py > gpurun_out/boxfit_$T.txt 2>&1; tail -3 gpurun_out/boxfit_$T.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$T.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_l_$T.log 2>&1; echo "ncu list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:scan_planes_kernel -s 4 -c 1 -o gpurun_out/prof_${T}_a_scan python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_a_$T.log 2>&1; echo "ncu scan rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:select_post_kernel -s 4 -c 1 -o gpurun_out/prof_${T}_b_post python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_b_$T.log 2>&1; echo "ncu post rc=$?"
