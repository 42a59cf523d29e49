mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -12) > gpurun_out/s4y_tests.log; cat gpurun_out/s4y_tests.log
cd /tmp && g++ -O2 -mavx -ffp-contract=off -w -DPANSLBM_B200_DROPIN -I$GRAFT_REPO_ROOT/include -I$GRAFT_REPO_ROOT/panslbm2_b200/src $GRAFT_REPO_ROOT/tests/dropin/ncpump_dump.cpp -o ncp -L$GRAFT_REPO_ROOT/panslbm2_b200 -lpanslbm_b200 -Wl,-rpath,$GRAFT_REPO_ROOT/panslbm2_b200 && mkdir -p o && ./ncp 51 101 5000 o
