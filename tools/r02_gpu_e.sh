#!/bin/bash
# round 2, GPU run E: L2 prefetch distance sweep at both block sizes, transient probe, drop-in / transient tests after the snapshot change
mkdir -p gpurun_out
O=gpurun_out
(timeout 600 python -m pytest tests/test_gpu_dropin.py tests/test_gpu_transient.py tests/test_gpu_closure_values.py tests/test_gpu_inplace.py -m gpu -q --maxfail=10 --timeout=300 2>&1 | tail -40) > $O/r02e_tests.log
B="timeout 300 python bench.py --steps 20 --warmup 5"
Q="--ns-size 0 --filter-size 0 --no-small --no-cpu"
for a in 18 37 74 111 148 185 222; do
  PANSLBM_L2_AHEAD=$a $B $Q > $O/r02e_bench_n1_ahead$a.json 2> $O/r02e_a$a.err
  PANSLBM_L2_AHEAD=$a $B --dims 81,161,81 $Q > $O/r02e_bench_81x161x81_ahead$a.json 2> $O/r02e_81a$a.err
done
PANSLBM_INPLACE=0 $B $Q > $O/r02e_bench_n1_two_buffers.json 2> $O/r02e_tb.err
for a in 0 74 148 296; do PANSLBM_L2_AHEAD=$a $B --size 32 --ns-size 512 --filter-size 0 --no-cpu > $O/r02e_bench_ns_ahead$a.json 2> $O/r02e_ns$a.err; done
PANSLBM_B200_PROFILE=1 timeout 400 python tools/transient_probe.py 200 > $O/r02e_transient_81x161x81_nt200.json 2> $O/r02e_transient.err
tail -12 $O/r02e_tests.log; cat $O/r02e_transient_81x161x81_nt200.json
python - <<'P'
import json, glob
for f in sorted(glob.glob("gpurun_out/r02e_bench*.json")):
    try:
        d = json.load(open(f))
        print(f.split("/")[-1], "value", round(d["value"], 1), "fwd/adj", round(d.get("sweeps", {}).get("forward_mlups", 0)), round(d.get("sweeps", {}).get("adjoint_mlups", 0)),
              "frac", round((d.get("roofline") or {}).get("frac", 0), 3), round((d.get("roofline_adjoint") or {}).get("frac", 0), 3), "ns", round((d.get("sweeps", {}).get("ns_cavity") or {}).get("mlups", 0)))
    except Exception as e:
        print(f, "FAILED", e)
P
