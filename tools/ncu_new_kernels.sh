mkdir -p gpurun_out
timeout 500 ncu --set full --clock-control none --import-source on -k regex:dst_blocked_gated -c 1 -f -o gpurun_out/r2_gated_8m python tools/microbench.py --uniform 8388608:262144:474 --dim 4096 --sum max --iters 1 > gpurun_out/ncu_gated.log 2>&1
timeout 500 ncu --set full --clock-control none --import-source on -k regex:weight_grad_tc -c 1 -f -o gpurun_out/r2_weight_grad_tc python tools/finetune_profile.py codex_l 64 > gpurun_out/ncu_wg.log 2>&1
timeout 500 ncu --set full --clock-control none --import-source on -k regex:dst_blocked_kernel -c 1 -f -o gpurun_out/r2_dst_blocked_16m_hints python tools/microbench.py --uniform 16777216:524288:474 --dim 4096 --iters 1 > gpurun_out/ncu_blocked.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -5
