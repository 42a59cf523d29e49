mkdir -p gpurun_out
run() { # name, env...
  name=$1; shift
  env "$@" timeout 200 python bench.py --no-cpu --ns-size 0 --steps 20 > gpurun_out/s4e_352_$name.json 2>gpurun_out/s4e_err.log
  env "$@" timeout 200 python bench.py --no-cpu --ns-size 0 --steps 200 --dims 81,161,81 > gpurun_out/s4e_prod_$name.json 2>>gpurun_out/s4e_err.log
}
run base A=1
run serial PANSLBM_SHELL_SERIAL=1
run serial_x1 PANSLBM_SHELL_SERIAL=1 PANSLBM_XINLINE=1
run blk64 PANSLBM_SHELL_BLOCK=64
run blk32 PANSLBM_SHELL_BLOCK=32
run ffirst PANSLBM_FUSED_FIRST=1
run slab8 PANSLBM_XSLAB=8
python - <<'P'
import json,glob
for f in sorted(glob.glob('gpurun_out/s4e_*.json')):
    try:
        d=json.load(open(f)); print(f, round(d['value']), round(d['sweeps']['forward_mlups']), round(d['sweeps']['adjoint_mlups']), round(d['ms_per_step'],4), round(d['roofline']['frac'],3), round(d['roofline']['avg_kernel_ms'],4), round(d['roofline']['kernel_share_of_timed_region'],3), round(d['e2e']['value']))
    except Exception as e: print(f, e)
P
timeout 300 python tools/e2e_probe.py 352 > gpurun_out/s4e_probe.log 2>&1; cat gpurun_out/s4e_probe.log
