mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6) > gpurun_out/s4t_tests.log; cat gpurun_out/s4t_tests.log
run() { # name, env...
  name=$1; shift
  env "$@" timeout 200 python bench.py --no-cpu --ns-size 0 --steps 20 > gpurun_out/s4t_352_$name.json 2>gpurun_out/s4t_err.log
  env "$@" timeout 200 python bench.py --no-cpu --ns-size 0 --steps 200 --dims 81,161,81 > gpurun_out/s4t_prod_$name.json 2>>gpurun_out/s4t_err.log
}
run base A=1
run graph PANSLBM_GRAPH=1
python - <<'P'
import json,glob
for f in sorted(glob.glob('gpurun_out/s4t_*.json')):
    try:
        d=json.load(open(f)); print(f, round(d['value']), round(d['sweeps']['forward_mlups']), round(d['sweeps']['adjoint_mlups']), round(d['ms_per_step'],4), round(d['roofline']['frac'],3), round(d['roofline']['avg_kernel_ms'],4), round(d['roofline']['kernel_share_of_timed_region'],3), round(d['e2e']['value']))
    except Exception as e: print(f, e)
P
for g in 0 1; do
  python tools/dropin_probe.py 2 141 161 1 5000 PANSLBM_GRAPH=$g
  python tools/dropin_probe.py 3 81 161 81 1000 PANSLBM_GRAPH=$g
done 2>&1 | tee gpurun_out/s4t_probe.log
