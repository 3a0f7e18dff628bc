timeout 500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for wl in cfg4 cfg2x cfg5; do
timeout 200 python bench.py --workload $wl --no-cpu-baseline --no-e2e --steps 100 > gpurun_out/pdl.json 2> gpurun_out/pdl.err || tail -5 gpurun_out/pdl.err
python -c "
import json; d=json.load(open('gpurun_out/pdl.json')); print('$wl', d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['step_frac'], d['roofline'].get('cold_ms'), d['roofline'].get('shift_ms_per_step'))"
done
RTM3D_B200_LIB=rtm3d_b200/librtm3d_decode_dev.so timeout 120 python tools/scan_stats.py cfg4 2>&1 | tail -4
