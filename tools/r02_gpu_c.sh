#!/bin/bash
# round 2, GPU run C: software-pipelined interior kernel (cp.async), in-place passes at 2 CTAs/SM, filters from patterns, closure values
mkdir -p gpurun_out
O=gpurun_out
(timeout 1800 python -m pytest tests -m gpu -q --maxfail=12 2>&1 | tail -120) > $O/r02c_tests.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 > $O/r02c_smoke.log
B="timeout 600 python bench.py --steps 20 --warmup 5"
$B > $O/r02c_bench_n1.json 2> $O/r02c_bench_n1.err
PANSLBM_PIPE=0 $B --no-cpu > $O/r02c_bench_n1_nopipe.json 2> $O/r02c_bench_np.err
PANSLBM_PIPE=0 PANSLBM_INPLACE=0 $B --ns-size 0 --no-cpu > $O/r02c_bench_n1_nopipe_two_buffers.json 2> $O/r02c_bench_nptb.err
PANSLBM_INPLACE=0 $B --ns-size 0 --no-cpu > $O/r02c_bench_n1_two_buffers.json 2> $O/r02c_bench_tb.err
PANSLBM_PIPE=0 PANSLBM_LIB_TAG=occ5 $B --ns-size 0 --no-cpu > $O/r02c_bench_n1_occ5_nopipe.json 2> $O/r02c_bench_occ5.err
$B --save-every-step --ns-size 0 --no-cpu > $O/r02c_bench_n1_save_every_step.json 2> $O/r02c_bench_ses.err
$B --dims 81,161,81 --ns-size 0 --no-cpu > $O/r02c_bench_81x161x81.json 2> $O/r02c_bench_81.err
PANSLBM_PIPE=0 $B --dims 81,161,81 --ns-size 0 --no-cpu > $O/r02c_bench_81x161x81_nopipe.json 2> $O/r02c_bench_81np.err
PANSLBM_GRAPH=1 $B --dims 81,161,81 --ns-size 0 --no-cpu > $O/r02c_bench_81x161x81_graph.json 2> $O/r02c_bench_81g.err
$B --size 512 --ns-size 0 --no-cpu > $O/r02c_bench_512.json 2> $O/r02c_bench_512.err
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > $O/r02c_bench_reference.json 2> $O/r02c_bench_reference.err
PANSLBM_B200_PROFILE=1 timeout 600 python tools/transient_probe.py 200 > $O/r02c_transient_81x161x81_nt200.json 2> $O/r02c_transient.err
NCU="ncu --clock-control none"
timeout 600 $NCU --metrics gpu__time_duration.sum -c 400 --csv --log-file $O/r02c_launches_bench_default.csv python bench.py --steps 4 --warmup 3 --ns-size 0 --no-cpu > $O/r02c_ncu_list.log 2>&1
timeout 600 $NCU --metrics gpu__time_duration.sum -c 400 --csv --log-file $O/r02c_launches_bench_81x161x81.csv python bench.py --steps 4 --warmup 3 --dims 81,161,81 --ns-size 0 --no-cpu > $O/r02c_ncu_list81.log 2>&1
# k_fused launch order with --steps 4 --warmup 3 (save_last = 2): fwd  S S | E E S S   adj  S S | E E S S ; passes alternate gather / local
for spec in "fwd_gather 2" "fwd_local 3" "adj_gather 8" "adj_local 9"; do
  set -- $spec
  timeout 900 $NCU --set full --import-source on -k regex:k_fused --launch-skip $2 --launch-count 1 -f -o /tmp/r02c_$1 python bench.py --steps 4 --warmup 3 --ns-size 0 --no-cpu > $O/r02c_ncu_full_$1.log 2>&1
  ncu -i /tmp/r02c_$1.ncu-rep --page raw --csv > $O/r02c_ncu_full_fused_$1_raw.csv 2>/dev/null
done
PANSLBM_PIPE=0 timeout 900 $NCU --set full -k regex:k_fused --launch-skip 3 --launch-count 1 -f -o /tmp/r02c_np python bench.py --steps 4 --warmup 3 --ns-size 0 --no-cpu > $O/r02c_ncu_full_nopipe.log 2>&1
ncu -i /tmp/r02c_np.ncu-rep --page raw --csv > $O/r02c_ncu_full_fused_fwd_local_nopipe_raw.csv 2>/dev/null
timeout 900 $NCU --set full --kernel-name-base mangled -k regex:k_fused.*ILi3ELi1E --launch-skip 2 --launch-count 2 -f -o /tmp/r02c_ns python bench.py --steps 4 --warmup 3 --size 32 --ns-size 512 --no-cpu > $O/r02c_ncu_full_ns.log 2>&1
ncu -i /tmp/r02c_ns.ncu-rep --page raw --csv > $O/r02c_ncu_full_fused_ns_raw.csv 2>/dev/null
timeout 600 $NCU --set full -k regex:k_xclose --launch-skip 3 --launch-count 1 -f -o /tmp/r02c_xc python bench.py --steps 4 --warmup 3 --ns-size 0 --no-cpu > $O/r02c_ncu_full_xclose.log 2>&1
ncu -i /tmp/r02c_xc.ncu-rep --page raw --csv > $O/r02c_ncu_full_xclose_raw.csv 2>/dev/null
tail -60 $O/r02c_tests.log; cat $O/r02c_smoke.log; cat $O/r02c_transient_81x161x81_nt200.json
python - <<'P'
import json, glob
for f in sorted(glob.glob("gpurun_out/r02c_bench*.json")):
    try:
        d = json.load(open(f))
        print(f.split("/")[-1], "value", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "fwd/adj", round(d.get("sweeps", {}).get("forward_mlups", 0)), round(d.get("sweeps", {}).get("adjoint_mlups", 0)),
              "frac", round((d.get("roofline") or {}).get("frac", 0), 3), round((d.get("roofline_adjoint") or {}).get("frac", 0), 3), "ns", round((d.get("sweeps", {}).get("ns_cavity") or {}).get("mlups", 0)))
    except Exception as e:
        print(f, "FAILED", e)
P
