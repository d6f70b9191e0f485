#!/bin/bash
# One-box multi-GPU measurements (run under `gpurun --gpus 8`): BASELINE.json configs[3] (60 s x B=256 sharded over 1/2/4/8 GPUs,
# strong scaling, bpf identical to the 1-GPU value) and configs[4] (DAC_VRVQ.encode end to end on 8 GPUs).  Results: gpurun_out/r2_*.json
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
for n in 1 2 4 8; do
  timeout 300 $TR --nproc-per-node $n --master-port $((29600 + n)) scripts/bench_cfg4_sharded.py --steps 8 --warmup 3 --out gpurun_out/r2_cfg4_${n}gpu.json 2> gpurun_out/r2_cfg4_${n}gpu.err | tail -1
done
# frame-split shards: B = 250 does not divide by 8 ranks in whole items
timeout 300 $TR --nproc-per-node 8 --master-port 29650 scripts/bench_cfg4_sharded.py --B 250 --steps 8 --warmup 3 --out gpurun_out/r2_cfg4_B250_8gpu.json 2>> gpurun_out/r2_cfg4_8gpu.err | tail -1
timeout 300 $TR --nproc-per-node 1 --master-port 29651 scripts/bench_cfg4_sharded.py --B 250 --steps 4 --warmup 2 --out gpurun_out/r2_cfg4_B250_1gpu.json 2>> gpurun_out/r2_cfg4_1gpu.err | tail -1
python - <<'PY'
import json
r = {n: json.load(open(f"gpurun_out/r2_cfg4_{n}gpu.json")) for n in (1, 2, 4, 8)}
base = r[1]
out = {"workload": "cfg4 sharded by sharding.plan_shards", "strong_scaling": {}}
for n, d in r.items():
    out["strong_scaling"][n] = {"frames_per_s": d["value"], "ms_per_pass": d["ms_per_pass_max_over_ranks"], "efficiency_vs_1gpu": d["value"] / (n * base["value"]),
                                "kept_equal_to_1gpu": d["kept_frames"] == base["kept_frames"], "bpf": d["bpf"], "bpf_equal_to_1gpu": d["bpf"] == base["bpf"]}
a, b = json.load(open("gpurun_out/r2_cfg4_B250_8gpu.json")), json.load(open("gpurun_out/r2_cfg4_B250_1gpu.json"))
out["frame_split_B250"] = {"frames_per_s_8gpu": a["value"], "kept_equal_to_1gpu": a["kept_frames"] == b["kept_frames"], "launches_per_pass_rank0": a["launches_per_pass_rank0"]}
json.dump(out, open("gpurun_out/r2_cfg4_scaling_summary.json", "w"), indent=1)
print(json.dumps(out))
PY
for n in 8 1; do
  timeout 300 $TR --nproc-per-node $n --master-port $((29700 + n)) scripts/bench_cfg5_encode.py --steps 20 --warmup 3 --out gpurun_out/r2_cfg5_${n}gpu.json 2> gpurun_out/r2_cfg5_${n}gpu.err | tail -1
done
timeout 300 $TR --nproc-per-node 8 --master-port 29720 scripts/bench_cfg5_encode.py --steps 20 --warmup 3 --tf32-encoder --out gpurun_out/r2_cfg5_8gpu_tf32enc.json 2>> gpurun_out/r2_cfg5_8gpu.err | tail -1
tail -3 gpurun_out/*.err | tail -30
