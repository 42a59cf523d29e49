#!/bin/bash
# round 2, GPU run A: full GPU suite, smoke, bench (default / production block / store-every-step A/B / reference arm),
# launch list and ncu --set full captures of the passes that store on the closure planes only
mkdir -p gpurun_out
O=gpurun_out
(timeout 1200 python -m pytest tests -m gpu -q --maxfail=25 2>&1 | tail -60) > $O/r02a_tests.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 > $O/r02a_smoke.log
timeout 600 python bench.py --steps 20 --warmup 5 > $O/r02a_bench_n1.json 2> $O/r02a_bench_n1.err
timeout 300 python bench.py --steps 20 --warmup 5 --dims 81,161,81 --ns-size 0 --no-cpu > $O/r02a_bench_81x161x81.json 2> $O/r02a_bench_81.err
timeout 300 python bench.py --steps 20 --warmup 5 --save-every-step --ns-size 0 --no-cpu > $O/r02a_bench_n1_save_every_step.json 2> $O/r02a_bench_ses.err
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > $O/r02a_bench_reference.json 2> $O/r02a_bench_reference.err
PANSLBM_B200_PROFILE=1 timeout 600 python tools/transient_probe.py 200 > $O/r02a_transient_81x161x81_nt200.json 2> $O/r02a_transient.err
# launch list (cold-cache, serialised: the SHARES count)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02a_launches_bench_default.csv python bench.py --steps 4 --warmup 3 --ns-size 0 --no-cpu > $O/r02a_ncu_list.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02a_launches_bench_81x161x81.csv python bench.py --steps 4 --warmup 3 --dims 81,161,81 --ns-size 0 --no-cpu > $O/r02a_ncu_list81.log 2>&1
# ncu --set full of one elided launch each (k_fused launch order with --steps 4 --warmup 3: fwd S S | E E S S, adj S S | E E S S)
for spec in "fwd 2" "adj 8"; do
  set -- $spec
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_fused --launch-skip $2 --launch-count 1 -f -o $O/r02a_ncu_full_fused_$1_elided python bench.py --steps 4 --warmup 3 --ns-size 0 --no-cpu > $O/r02a_ncu_full_$1.log 2>&1
  ncu -i $O/r02a_ncu_full_fused_$1_elided.ncu-rep --page raw --csv > $O/r02a_ncu_full_fused_$1_elided_raw.csv 2>/dev/null
done
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:k_fusedILi3ELi1E --launch-skip 2 --launch-count 1 -f -o $O/r02a_ncu_full_fused_ns_elided python bench.py --steps 4 --warmup 3 --size 32 --ns-size 512 --no-cpu > $O/r02a_ncu_full_ns.log 2>&1
ncu -i $O/r02a_ncu_full_fused_ns_elided.ncu-rep --page raw --csv > $O/r02a_ncu_full_fused_ns_elided_raw.csv 2>/dev/null
timeout 600 ncu --set full --clock-control none -k regex:k_xclose --launch-skip 3 --launch-count 1 -f -o $O/r02a_ncu_full_xclose_81 python bench.py --steps 4 --warmup 3 --dims 81,161,81 --ns-size 0 --no-cpu > $O/r02a_ncu_full_xclose81.log 2>&1
ncu -i $O/r02a_ncu_full_xclose_81.ncu-rep --page raw --csv > $O/r02a_ncu_full_xclose_81_raw.csv 2>/dev/null
rm -f $O/*.ncu-rep.tmp
ls -la $O | tail -30
tail -3 $O/r02a_tests.log; cat $O/r02a_smoke.log; cat $O/r02a_transient_81x161x81_nt200.json; grep "host profile" $O/r02a_transient.err | tail -30
python - <<'P'
import json
for f in ("r02a_bench_n1", "r02a_bench_81x161x81", "r02a_bench_n1_save_every_step", "r02a_bench_reference"):
    try:
        d = json.load(open(f"gpurun_out/{f}.json"))
        print(f, "value", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "fwd/adj", d.get("sweeps", {}).get("forward_mlups"), d.get("sweeps", {}).get("adjoint_mlups"),
              "frac", (d.get("roofline") or {}).get("frac"), (d.get("roofline_adjoint") or {}).get("frac"), "ns", (d.get("sweeps", {}).get("ns_cavity") or {}).get("mlups"))
    except Exception as e:
        print(f, "FAILED", e)
P
