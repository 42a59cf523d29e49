#!/bin/bash
# Round 2, final single-GPU run: the whole GPU suite, the default bench line, the production block, the checkpoint probe.
# Everything lands in gpurun_out/r02n_* (tools/collect_profiles.py / by hand into profiles/).
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
python -m pytest tests -m gpu -q --timeout 300 > $O/r02n_tests.log 2>&1
tail -25 $O/r02n_tests.log | cut -c1-400
python bench.py > $O/r02n_bench_n1.json 2> $O/r02n_bench_n1.err; tail -c 600 $O/r02n_bench_n1.json; tail -3 $O/r02n_bench_n1.err
python bench.py --dims 81,161,81 > $O/r02n_bench_81x161x81.json 2> $O/r02n_bench_81.err; tail -c 300 $O/r02n_bench_81x161x81.json
python tools/checkpoint_probe.py --every 1,14 > $O/r02n_checkpoint_probe2.json 2> $O/r02n_checkpoint_probe2.err; cat $O/r02n_checkpoint_probe2.json
