# bench.py on N GPUs of one box (torchrun), output to gpurun_out/final_bench_n$N.json:  bash tools/multi_bench.sh 8
mkdir -p gpurun_out
N=$1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/final_bench_n$N.json 2> gpurun_out/final_bench_n$N.err
tail -c 400 gpurun_out/final_bench_n$N.json; tail -3 gpurun_out/final_bench_n$N.err
