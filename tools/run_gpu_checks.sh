#!/bin/bash
# What a round checks on a B200 box (under gpurun): the GPU suite, the smoke entry, our bench arm and the reference arm.
#   gpurun --timeout 1500 -- 'bash tools/run_gpu_checks.sh'        outputs under gpurun_out/
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6) | tee gpurun_out/check_tests.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/check_smoke.log
timeout 400 python bench.py > gpurun_out/check_bench.json 2> gpurun_out/check_bench.err
timeout 400 python bench.py --impl reference > gpurun_out/check_bench_reference.json 2> gpurun_out/check_bench_reference.err
python - <<'P'
import json
for f in ("gpurun_out/check_bench.json", "gpurun_out/check_bench_reference.json"):
    d = json.load(open(f))
    print(f, "value", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "roofline.frac", (d.get("roofline") or {}).get("frac"))
P
