#!/bin/bash
# round 2, GPU run H: cooperative multi-step launch (k_steps): full GPU suite, small-domain sub-lines with / without it, production block, default bench
mkdir -p gpurun_out
O=gpurun_out
(timeout 1200 python -m pytest tests -m gpu -q --maxfail=10 --timeout=400 2>&1 | tail -80) > $O/r02h_tests.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 > $O/r02h_smoke.log
B="timeout 500 python bench.py --steps 20 --warmup 5"
$B > $O/r02h_bench_n1.json 2> $O/r02h_bench_n1.err
PANSLBM_COOP_SITES=0 $B --ns-size 0 --filter-size 0 > $O/r02h_bench_n1_nocoop.json 2> $O/r02h_bench_nocoop.err
$B --dims 81,161,81 --ns-size 0 --filter-size 0 --no-small --no-cpu > $O/r02h_bench_81x161x81.json 2> $O/r02h_bench_81.err
$B --dims 41,81,41 --ns-size 0 --filter-size 0 --no-small --no-cpu > $O/r02h_bench_41x81x41.json 2> $O/r02h_bench_41.err
PANSLBM_COOP_SITES=0 $B --dims 41,81,41 --ns-size 0 --filter-size 0 --no-small --no-cpu > $O/r02h_bench_41x81x41_nocoop.json 2> $O/r02h_bench_41nc.err
for i in 1 2; do PANSLBM_B200_PROFILE=1 timeout 400 python tools/transient_probe.py 200 > $O/r02h_transient_81x161x81_nt200_$i.json 2> $O/r02h_transient_$i.err; done
# the unmodified 2-D driver loop through the drop-in engine (heatsink_dump at the committed size of production/heatsink.cpp)
tail -30 $O/r02h_tests.log; cat $O/r02h_smoke.log; cat $O/r02h_transient_81x161x81_nt200_1.json; cat $O/r02h_transient_81x161x81_nt200_2.json
python - <<'P'
import json, glob
for f in sorted(glob.glob("gpurun_out/r02h_bench*.json")):
    try:
        d = json.load(open(f))
        print(f.split("/")[-1], "value", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "fwd/adj", round(d["sweeps"]["forward_mlups"]), round(d["sweeps"]["adjoint_mlups"]), "launches", d["gpu_launches"])
        for k, v in (d["sweeps"].get("small_domains") or {}).items():
            print("    ", k, {a: (round(b, 2) if isinstance(b, float) else b) for a, b in v.items() if a != "workload"} if isinstance(v, dict) else v)
        if "filter" in d["sweeps"]: print("    filter", d["sweeps"]["filter"])
    except Exception as e:
        print(f, "FAILED", e)
P
