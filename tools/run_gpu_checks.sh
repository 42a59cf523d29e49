mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 400 python bench.py > gpurun_out/r01_final_bench_n1.json 2> gpurun_out/r01_final_bench_n1.err; tail -c 400 gpurun_out/r01_final_bench_n1.json; echo
timeout 400 python bench.py --impl reference > gpurun_out/r01_final_bench_reference.json 2> gpurun_out/r01_final_bench_reference.err; tail -c 700 gpurun_out/r01_final_bench_reference.json; echo
timeout 300 python bench.py --no-cpu --ns-size 0 --steps 200 --dims 81,161,81 > gpurun_out/r01_final_bench_81x161x81.json 2>/dev/null
python tools/transient_probe.py 200 0 20 > gpurun_out/r01_final_transient.json 2>/dev/null; cat gpurun_out/r01_final_transient.json
