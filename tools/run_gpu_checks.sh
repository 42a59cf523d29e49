mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6) > gpurun_out/s4w_tests.log; cat gpurun_out/s4w_tests.log
timeout 200 python bench.py --no-cpu --ns-size 0 --steps 20 > gpurun_out/s4w_352.json 2>gpurun_out/s4w_err.log
timeout 200 python bench.py --no-cpu --ns-size 0 --steps 200 --dims 81,161,81 > gpurun_out/s4w_prod.json 2>>gpurun_out/s4w_err.log
python - <<'P'
import json,glob
for f in sorted(glob.glob('gpurun_out/s4w_*.json')):
    try:
        d=json.load(open(f)); print(f, round(d['value']), round(d['sweeps']['forward_mlups']), round(d['sweeps']['adjoint_mlups']), round(d['ms_per_step'],4), round(d['roofline']['frac'],3), round(d['roofline']['avg_kernel_ms'],4), round(d['roofline']['kernel_share_of_timed_region'],3), round(d['e2e']['value']))
    except Exception as e: print(f, e)
P
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_xclose -c 12 --csv python bench.py --no-cpu --ns-size 0 --steps 3 2>/dev/null | grep k_xclose | awk -F, '{print $(NF)}' | tr '\n' ' '
