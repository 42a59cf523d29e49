#!/bin/bash
# Round 2: schedule knobs in the launch-bound regime (2-D domains, small 3-D blocks), one process per setting.
cd "$(dirname "$0")/.."
O=gpurun_out/r02o_small_domain_knobs.jsonl
: > $O
run() { env "$@" python tools/small_domain_probe.py >> $O 2>> gpurun_out/r02o.err; }
run PANSLBM_NOP=1
run PANSLBM_GRAPH=1
run PANSLBM_XGHOST=0
run PANSLBM_XGHOST=0 PANSLBM_GRAPH=1
run PANSLBM_SHELL_SERIAL=1
run PANSLBM_XGHOST=0 PANSLBM_SHELL_SERIAL=1
run PANSLBM_XGHOST=0 PANSLBM_XINLINE=1 PANSLBM_GRAPH=1
run PANSLBM_FUSED_FIRST=1 PANSLBM_GRAPH=1
run PANSLBM_L2_AHEAD=0 PANSLBM_GRAPH=1
cat $O | cut -c1-900
tail -5 gpurun_out/r02o.err
