#!/bin/bash
# round 2, GPU run J: final code (straight-line wall_scatter, filter ghost exchange): full GPU suite + the bench lines that depend on the interior kernel
mkdir -p gpurun_out
O=gpurun_out
(timeout 1200 python -m pytest tests -m gpu -q --maxfail=10 --timeout=400 2>&1 | tail -40) > $O/r02j_tests.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 > $O/r02j_smoke.log
B="timeout 500 python bench.py --steps 20 --warmup 5"
Q="--ns-size 0 --filter-size 0 --no-small --no-cpu"
$B > $O/r02j_bench_n1.json 2> $O/r02j_bench_n1.err
PANSLBM_INPLACE=0 $B $Q > $O/r02j_bench_n1_two_buffers.json 2> $O/r02j_tb.err
$B --save-every-step $Q > $O/r02j_bench_n1_save_every_step.json 2> $O/r02j_ses.err
$B --dims 81,161,81 $Q > $O/r02j_bench_81x161x81.json 2> $O/r02j_81.err
$B --dims 41,81,41 $Q > $O/r02j_bench_41x81x41.json 2> $O/r02j_41.err
$B --size 512 $Q > $O/r02j_bench_512.json 2> $O/r02j_512.err
NCU="ncu --clock-control none"
timeout 400 $NCU --metrics gpu__time_duration.sum -c 400 --csv --log-file $O/r02j_launches_bench_default.csv python bench.py --steps 4 --warmup 3 $Q > $O/r02j_ncu_list.log 2>&1
timeout 400 $NCU --metrics gpu__time_duration.sum -c 400 --csv --log-file $O/r02j_launches_bench_81x161x81.csv python bench.py --steps 4 --warmup 3 --dims 81,161,81 $Q > $O/r02j_ncu_list81.log 2>&1
for spec in "fwd_gather 2" "fwd_local 3" "adj_gather 8" "adj_local 9"; do
  set -- $spec
  timeout 600 $NCU --set full --import-source on -k regex:k_fused --launch-skip $2 --launch-count 1 -f -o /tmp/r02j_$1 python bench.py --steps 4 --warmup 3 $Q > $O/r02j_ncu_full_$1.log 2>&1
  ncu -i /tmp/r02j_$1.ncu-rep --page raw --csv > $O/r02j_ncu_full_fused_$1_raw.csv 2>/dev/null
done
tail -6 $O/r02j_tests.log; cat $O/r02j_smoke.log
python - <<'P'
import json, glob
for f in sorted(glob.glob("gpurun_out/r02j_bench*.json")):
    try:
        d = json.load(open(f))
        print(f.split("/")[-1], "value", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "fwd/adj", round(d["sweeps"]["forward_mlups"]), round(d["sweeps"]["adjoint_mlups"]),
              "frac", round((d.get("roofline") or {}).get("frac", 0), 3), round((d.get("roofline_adjoint") or {}).get("frac", 0), 3), "ns", round((d["sweeps"].get("ns_cavity") or {}).get("mlups", 0)))
        if "filter" in d["sweeps"]: print("    filter", {a: b for a, b in d["sweeps"]["filter"].items() if a != "workload"})
    except Exception as e:
        print(f, "FAILED", e)
P
