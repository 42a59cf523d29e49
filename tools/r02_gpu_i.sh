#!/bin/bash
# round 2, final 1-GPU run: full GPU suite, smoke, every bench line kept under profiles/, launch lists, ncu --set full captures (CSV only)
mkdir -p gpurun_out
O=gpurun_out
(timeout 1200 python -m pytest tests -m gpu -q --maxfail=10 --timeout=400 2>&1 | tail -40) > $O/r02i_tests.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 > $O/r02i_smoke.log
B="timeout 500 python bench.py --steps 20 --warmup 5"
Q="--ns-size 0 --filter-size 0 --no-small --no-cpu"
$B > $O/r02i_bench_n1.json 2> $O/r02i_bench_n1.err
timeout 500 python bench.py --impl reference --steps 20 --warmup 5 > $O/r02i_bench_reference.json 2> $O/r02i_bench_reference.err
PANSLBM_INPLACE=0 $B $Q > $O/r02i_bench_n1_two_buffers.json 2> $O/r02i_tb.err
$B --save-every-step $Q > $O/r02i_bench_n1_save_every_step.json 2> $O/r02i_ses.err
$B --dims 81,161,81 $Q > $O/r02i_bench_81x161x81.json 2> $O/r02i_81.err
$B --size 512 $Q > $O/r02i_bench_512.json 2> $O/r02i_512.err
for i in 1 2; do timeout 400 python tools/transient_probe.py 200 > $O/r02i_transient_81x161x81_nt200_$i.json 2> $O/r02i_transient_$i.err; done
PANSLBM_B200_DEVICE_BUDGET_MB=8000 PROBE_ITERATIONS=1 timeout 600 python tools/transient_probe.py 200 > $O/r02i_transient_81x161x81_nt200_budget8GB.json 2> $O/r02i_transient_b.err
NCU="ncu --clock-control none"
timeout 400 $NCU --metrics gpu__time_duration.sum -c 400 --csv --log-file $O/r02i_launches_bench_default.csv python bench.py --steps 4 --warmup 3 $Q > $O/r02i_ncu_list.log 2>&1
timeout 400 $NCU --metrics gpu__time_duration.sum -c 400 --csv --log-file $O/r02i_launches_bench_81x161x81.csv python bench.py --steps 4 --warmup 3 --dims 81,161,81 $Q > $O/r02i_ncu_list81.log 2>&1
# k_fused launch order with --steps 4 --warmup 3 (save_last = 2): fwd  S S | E E S S   adj  S S | E E S S ; passes alternate gather / local
for spec in "fwd_gather 2" "fwd_local 3" "fwd_storing 4" "adj_gather 8" "adj_local 9" "adj_storing 10"; do
  set -- $spec
  timeout 600 $NCU --set full --import-source on -k regex:k_fused --launch-skip $2 --launch-count 1 -f -o /tmp/r02i_$1 python bench.py --steps 4 --warmup 3 $Q > $O/r02i_ncu_full_$1.log 2>&1
  ncu -i /tmp/r02i_$1.ncu-rep --page raw --csv > $O/r02i_ncu_full_fused_$1_raw.csv 2>/dev/null
done
timeout 600 $NCU --set full --kernel-name-base mangled -k regex:k_fusedILi3ELi1E --launch-skip 2 --launch-count 2 -f -o /tmp/r02i_ns python bench.py --steps 4 --warmup 3 --size 32 --ns-size 512 --filter-size 0 --no-small --no-cpu > $O/r02i_ncu_full_ns.log 2>&1
ncu -i /tmp/r02i_ns.ncu-rep --page raw --csv > $O/r02i_ncu_full_fused_ns_raw.csv 2>/dev/null
for k in k_xclose k_shell k_tubes k_sensitivity; do
  timeout 400 $NCU --set full -k regex:$k --launch-skip 2 --launch-count 1 -f -o /tmp/r02i_$k python bench.py --steps 4 --warmup 3 $Q > /dev/null 2>&1
  ncu -i /tmp/r02i_$k.ncu-rep --page raw --csv > $O/r02i_ncu_full_${k}_raw.csv 2>/dev/null
done
timeout 400 $NCU --set full -k regex:k_filter --launch-skip 4 --launch-count 1 -f -o /tmp/r02i_k_filter python bench.py --steps 4 --warmup 3 --size 32 --ns-size 0 --filter-size 128 --no-small --no-cpu > /dev/null 2>&1
ncu -i /tmp/r02i_k_filter.ncu-rep --page raw --csv > $O/r02i_ncu_full_k_filter_raw.csv 2>/dev/null
timeout 300 $NCU --set full -k regex:k_residual_partial --launch-count 1 -f -o /tmp/r02i_res python -c "
import sys; sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import panslbm2_b200 as pl
n = 352**3
a = [pl.DeviceArray(n, 0.5 + 0.1*k) for k in range(6)]
print(pl.Residual(*a, n))" > /dev/null 2>&1
ncu -i /tmp/r02i_res.ncu-rep --page raw --csv > $O/r02i_ncu_full_k_residual_raw.csv 2>/dev/null
tail -6 $O/r02i_tests.log; cat $O/r02i_smoke.log
python - <<'P'
import json, glob
for f in sorted(glob.glob("gpurun_out/r02i_bench*.json")) + sorted(glob.glob("gpurun_out/r02i_transient*.json")):
    try:
        d = json.load(open(f))
        if "value" in d:
            print(f.split("/")[-1], "value", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "fwd/adj", round(d.get("sweeps", {}).get("forward_mlups", 0)), round(d.get("sweeps", {}).get("adjoint_mlups", 0)),
                  "frac", round((d.get("roofline") or {}).get("frac", 0), 3), round((d.get("roofline_adjoint") or {}).get("frac", 0), 3), "ns", round((d.get("sweeps", {}).get("ns_cavity") or {}).get("mlups", 0)))
            for k, v in (d.get("sweeps", {}).get("small_domains") or {}).items():
                print("    ", k, {a: (round(b, 2) if isinstance(b, float) else b) for a, b in v.items() if a != "workload"} if isinstance(v, dict) else v)
            if "filter" in d.get("sweeps", {}): print("    filter", {a: b for a, b in d["sweeps"]["filter"].items() if a != "workload"})
        else:
            print(f.split("/")[-1], {k: (v.get("forward_ms_per_step"), v.get("adjoint_ms_per_step"), v.get("spilled")) for k, v in d.items() if isinstance(v, dict)})
    except Exception as e:
        print(f, "FAILED", e)
P
