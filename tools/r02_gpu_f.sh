#!/bin/bash
# round 2, GPU run F (2 GPUs): NCCL parity tests, 2-rank transient parity, 2-GPU unmodified heavisidefilter, bench N = 2 with parity, reference arm under torchrun
mkdir -p gpurun_out
O=gpurun_out
(timeout 900 python -m pytest tests/test_gpu_nccl.py tests/test_gpu_transient.py tests/test_gpu_dropin.py -m gpu -q -k "nccl or over_ranks or heavisidefilter" --maxfail=10 --timeout=400 2>&1 | tail -40) > $O/r02f_tests.log
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $T bench.py --gpus 2 --steps 20 --warmup 5 > $O/r02f_bench_n2.json 2> $O/r02f_bench_n2.err
timeout 600 $T bench.py --gpus 2 --steps 20 --warmup 5 --impl reference > $O/r02f_bench_n2_reference.json 2> $O/r02f_bench_n2_reference.err
timeout 300 $T bench.py --gpus 2 --steps 20 --warmup 5 --global-size 512 --ns-size 0 --filter-size 0 > $O/r02f_bench_n2_strong512.json 2> $O/r02f_bench_n2_strong.err
tail -15 $O/r02f_tests.log
python - <<'P'
import json, glob
for f in sorted(glob.glob("gpurun_out/r02f_bench*.json")):
    try:
        d = json.load(open(f))
        print(f.split("/")[-1], "value", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "scaling", d.get("scaling"), "parity", d.get("parity"), "cores", (d.get("cpu_baseline") or {}).get("cores"))
    except Exception as e:
        print(f, "FAILED", e, open(f.replace(".json", ".err")).read()[-1500:] if False else "")
P
tail -5 $O/r02f_bench_n2.err
