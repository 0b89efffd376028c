mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/r1_pytest.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 > gpurun_out/r1_smoke.txt
timeout 900 python bench.py > gpurun_out/r1_bench.json 2> gpurun_out/r1_bench.err
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum
timeout 300 ncu --profile-from-start off --metrics $M --clock-control none --csv --log-file gpurun_out/traffic_c2.csv python tools/one_step.py --graph fb15k237 --batch 64 > gpurun_out/traffic_c2.log 2>&1
timeout 300 ncu --profile-from-start off --metrics $M --clock-control none --csv --log-file gpurun_out/traffic_c4.csv python tools/one_step.py --graph yago310 --batch 64 > gpurun_out/traffic_c4.log 2>&1
timeout 600 ncu --profile-from-start off --metrics $M --clock-control none --csv --log-file gpurun_out/traffic_c5.csv python tools/one_step.py --uniform 16777216:524288:474 --dim 4096 > gpurun_out/traffic_c5.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:seg_reduce -c 3 -f -o gpurun_out/r2_c2_step python tools/one_step.py --graph fb15k237 --batch 64 > gpurun_out/ncu_c2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:linear_norm_relu_residual_tc -c 1 -f -o gpurun_out/r2_linear_ts python tools/linear_bench.py --iters 2 > gpurun_out/ncu_linear.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 3 --warmup 3 > gpurun_out/ncu_bench.log 2>&1
cat gpurun_out/r1_pytest.txt gpurun_out/r1_smoke.txt; tail -c 600 gpurun_out/r1_bench.json
