set -u
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
for n in 1 8 4 2; do
  timeout 200 $TR --nproc-per-node $n --master-port $((29600 + n)) scripts/bench_cfg4_sharded.py --steps 8 --warmup 3 --out gpurun_out/r2u_cfg4_${n}gpu.json 2> gpurun_out/r2u_cfg4_${n}gpu.err | tail -1
done
