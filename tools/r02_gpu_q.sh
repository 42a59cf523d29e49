#!/bin/bash
# Round 2, last single-GPU run of the final library: the whole GPU suite, smoke, and configs[4] with nt = 1000 stored steps
# (194 GB of states in the reference's scheme) through the checkpoint-recompute store.
cd "$(dirname "$0")/.."
O=gpurun_out
python -m pytest tests -m gpu -q --timeout 300 > $O/r02q_tests.log 2>&1
tail -8 $O/r02q_tests.log | cut -c1-400
python -c "import __graft_entry__ as g; g.smoke()" > $O/r02q_smoke.log 2>&1; tail -2 $O/r02q_smoke.log
python tools/checkpoint_probe.py --nt 1000 --every 32,50 > $O/r02q_checkpoint_probe_nt1000.json 2> $O/r02q_checkpoint_probe_nt1000.err; cat $O/r02q_checkpoint_probe_nt1000.json; tail -2 $O/r02q_checkpoint_probe_nt1000.err
