# e2e leg of bench.py at 4 GPUs with 128- and 512-column chunks of the host-buffer pipeline (ULTRA_RSPMM_CHUNK_COLS)
mkdir -p gpurun_out
for c in 128 512; do
ULTRA_RSPMM_CHUNK_COLS=$c timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 4 --steps 10 --warmup 3 > gpurun_out/chunk_$c.json 2> gpurun_out/chunk_$c.err
python - <<PY
import json
d=json.loads(open('gpurun_out/chunk_$c.json').read().strip().splitlines()[-1])
print("chunk $c e2e ms", d['e2e']['ms_per_step'], "ceiling", d['e2e']['host_copy_ceiling']['ms_per_step'])
PY
done
