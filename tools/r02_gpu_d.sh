#!/bin/bash
# round 2, GPU run D: full GPU suite (recursive block-map lock), L2 prefetch-ahead A/B, two-buffer A/B, production block, NS, ncu
mkdir -p gpurun_out
O=gpurun_out
(timeout 1200 python -m pytest tests -m gpu -q --maxfail=10 --timeout=400 2>&1 | tail -150) > $O/r02d_tests.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 > $O/r02d_smoke.log
B="timeout 400 python bench.py --steps 20 --warmup 5"
Q="--ns-size 0 --filter-size 0 --no-small --no-cpu"
$B > $O/r02d_bench_n1.json 2> $O/r02d_bench_n1.err
for a in 0 148 592 1184; do PANSLBM_L2_AHEAD=$a $B $Q > $O/r02d_bench_n1_ahead$a.json 2> $O/r02d_bench_a$a.err; done
PANSLBM_INPLACE=0 $B $Q > $O/r02d_bench_n1_two_buffers.json 2> $O/r02d_bench_tb.err
PANSLBM_INPLACE=0 PANSLBM_L2_AHEAD=0 $B $Q > $O/r02d_bench_n1_two_buffers_ahead0.json 2> $O/r02d_bench_tb0.err
$B --save-every-step $Q > $O/r02d_bench_n1_save_every_step.json 2> $O/r02d_bench_ses.err
$B --dims 81,161,81 $Q > $O/r02d_bench_81x161x81.json 2> $O/r02d_bench_81.err
PANSLBM_L2_AHEAD=0 $B --dims 81,161,81 $Q > $O/r02d_bench_81x161x81_ahead0.json 2> $O/r02d_bench_81a0.err
PANSLBM_GRAPH=1 $B --dims 81,161,81 $Q > $O/r02d_bench_81x161x81_graph.json 2> $O/r02d_bench_81g.err
$B --size 512 $Q > $O/r02d_bench_512.json 2> $O/r02d_bench_512.err
PANSLBM_L2_AHEAD=0 $B --size 32 --filter-size 0 --no-cpu > $O/r02d_bench_ns_ahead0.json 2> $O/r02d_bench_nsa0.err
timeout 400 python bench.py --impl reference --steps 20 --warmup 5 > $O/r02d_bench_reference.json 2> $O/r02d_bench_reference.err
PANSLBM_B200_PROFILE=1 timeout 400 python tools/transient_probe.py 200 > $O/r02d_transient_81x161x81_nt200.json 2> $O/r02d_transient.err
NCU="ncu --clock-control none"
timeout 400 $NCU --metrics gpu__time_duration.sum -c 400 --csv --log-file $O/r02d_launches_bench_default.csv python bench.py --steps 4 --warmup 3 $Q > $O/r02d_ncu_list.log 2>&1
timeout 400 $NCU --metrics gpu__time_duration.sum -c 400 --csv --log-file $O/r02d_launches_bench_81x161x81.csv python bench.py --steps 4 --warmup 3 --dims 81,161,81 $Q > $O/r02d_ncu_list81.log 2>&1
for spec in "fwd_gather 2" "fwd_local 3" "adj_gather 8" "adj_local 9"; do
  set -- $spec
  timeout 600 $NCU --set full --import-source on -k regex:k_fused --launch-skip $2 --launch-count 1 -f -o /tmp/r02d_$1 python bench.py --steps 4 --warmup 3 $Q > $O/r02d_ncu_full_$1.log 2>&1
  ncu -i /tmp/r02d_$1.ncu-rep --page raw --csv > $O/r02d_ncu_full_fused_$1_raw.csv 2>/dev/null
done
timeout 600 $NCU --set full --kernel-name-base mangled -k regex:k_fusedILi3ELi1E --launch-skip 2 --launch-count 2 -f -o /tmp/r02d_ns python bench.py --steps 4 --warmup 3 --size 32 --ns-size 512 --filter-size 0 --no-cpu > $O/r02d_ncu_full_ns.log 2>&1
ncu -i /tmp/r02d_ns.ncu-rep --page raw --csv > $O/r02d_ncu_full_fused_ns_raw.csv 2>/dev/null
for k in k_shell k_tubes k_filter k_sensitivity k_residual_partial; do
  timeout 300 $NCU --set full -k regex:$k --launch-skip 1 --launch-count 1 -f -o /tmp/r02d_$k python bench.py --steps 4 --warmup 3 --size 128 --ns-size 0 --filter-size 64 --no-cpu > /dev/null 2>&1
  ncu -i /tmp/r02d_$k.ncu-rep --page raw --csv > $O/r02d_ncu_full_${k}_raw.csv 2>/dev/null
done
tail -50 $O/r02d_tests.log; cat $O/r02d_smoke.log; cat $O/r02d_transient_81x161x81_nt200.json
python - <<'P'
import json, glob
for f in sorted(glob.glob("gpurun_out/r02d_bench*.json")):
    try:
        d = json.load(open(f))
        print(f.split("/")[-1], "value", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "fwd/adj", round(d.get("sweeps", {}).get("forward_mlups", 0)), round(d.get("sweeps", {}).get("adjoint_mlups", 0)),
              "frac", round((d.get("roofline") or {}).get("frac", 0), 3), round((d.get("roofline_adjoint") or {}).get("frac", 0), 3), "ns", round((d.get("sweeps", {}).get("ns_cavity") or {}).get("mlups", 0)),
              "filter", (d.get("sweeps", {}).get("filter") or {}).get("gpu_ms_per_call"), (d.get("sweeps", {}).get("filter") or {}).get("reference_ms_per_call"))
    except Exception as e:
        print(f, "FAILED", e)
P
