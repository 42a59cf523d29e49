mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3) > gpurun_out/s4l_tests.log; cat gpurun_out/s4l_tests.log
run() { # name, env...
  name=$1; shift
  env "$@" timeout 200 python bench.py --no-cpu --ns-size 0 --steps 20 > gpurun_out/s4l_352_$name.json 2>gpurun_out/s4l_err.log
  env "$@" timeout 200 python bench.py --no-cpu --ns-size 0 --steps 200 --dims 81,161,81 > gpurun_out/s4l_prod_$name.json 2>>gpurun_out/s4l_err.log
}
run p0 PANSLBM_PREFETCH=0
run p3 PANSLBM_PREFETCH=3
run p1 PANSLBM_PREFETCH=1
run p3_serial PANSLBM_PREFETCH=3 PANSLBM_SHELL_SERIAL=1
python - <<'P'
import json,glob
for f in sorted(glob.glob('gpurun_out/s4l_*.json')):
    try:
        d=json.load(open(f)); print(f, round(d['value']), round(d['sweeps']['forward_mlups']), round(d['sweeps']['adjoint_mlups']), round(d['ms_per_step'],4), round(d['roofline']['frac'],3), round(d['roofline']['avg_kernel_ms'],4), round(d['roofline']['kernel_share_of_timed_region'],3), round(d['e2e']['value']))
    except Exception as e: print(f, e)
P
