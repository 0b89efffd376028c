# Round-end evidence run on one B200 (gpurun): GPU tests, smoke, the bench line, the configs[4] sweep.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/final_pytest.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 > gpurun_out/final_smoke.txt
timeout 900 python bench.py > gpurun_out/final_bench_n1.json 2> gpurun_out/final_bench_n1.err
timeout 1500 python tools/sweep.py > gpurun_out/final_sweep.csv 2> gpurun_out/final_sweep.err
cat gpurun_out/final_pytest.txt gpurun_out/final_smoke.txt; tail -c 300 gpurun_out/final_bench_n1.json; wc -l gpurun_out/final_sweep.csv
