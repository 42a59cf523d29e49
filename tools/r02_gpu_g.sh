#!/bin/bash
# round 2, GPU run G (8 GPUs): configs[3] as written (81x161x81 on 2x2x2, golden parity), weak scaling 352^3 per GPU, strong scaling 512^3,
# 2x2x2 NCCL parity test, transient loops on 2x2x2 (parity + probe)
mkdir -p gpurun_out
O=gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512"
timeout 600 $T bench.py --gpus 8 --steps 20 --warmup 5 --config heatsink3d > $O/r02g_bench_n8_heatsink3d.json 2> $O/r02g_bench_n8_heatsink3d.err
timeout 600 $T bench.py --gpus 8 --steps 20 --warmup 5 > $O/r02g_bench_n8.json 2> $O/r02g_bench_n8.err
timeout 400 $T bench.py --gpus 8 --steps 20 --warmup 5 --global-size 512 --ns-size 0 --filter-size 0 --no-parity > $O/r02g_bench_n8_strong512.json 2> $O/r02g_bench_n8_strong.err
(timeout 600 python -m pytest tests/test_gpu_nccl.py tests/test_gpu_transient.py -m gpu -q -k "2x2x2 or 2,2,2" --maxfail=4 --timeout=400 2>&1 | tail -15) > $O/r02g_tests.log
PROBE_PE=2,2,2 PROBE_ITERATIONS=2 timeout 400 python tools/transient_probe.py 200 > $O/r02g_transient_81x161x81_nt200_pe222.json 2> $O/r02g_transient.err
tail -6 $O/r02g_tests.log; cat $O/r02g_transient_81x161x81_nt200_pe222.json
python - <<'P'
import json, glob
for f in sorted(glob.glob("gpurun_out/r02g_bench*.json")):
    try:
        d = json.load(open(f))
        print(f.split("/")[-1], "value", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "scaling", d.get("scaling"), "parity", d.get("parity"), d.get("parity_golden"))
    except Exception as e:
        print(f, "FAILED", e)
P
tail -4 $O/r02g_bench_n8_heatsink3d.err
