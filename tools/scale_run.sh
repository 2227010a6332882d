#!/bin/bash
# The round's multi-GPU measurements, one JSON line per run.  N GPUs cost N x the box time, so every rank count gets its
# own call:  gpurun --gpus 8 -- bash tools/scale_run.sh 8 ;  gpurun --gpus 4 -- bash tools/scale_run.sh 4 ; ...
n=${1:-8}
out=gpurun_out/r3_scale_n$n.jsonl
err=gpurun_out/r3_scale_n$n.err
: > $out; : > $err
tr() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $n "$@"; }
tr --steps 20 --warmup 5 --no-extras --shard database >> $out 2>> $err   # strong, index work sharded over NVLink peer memory
tr --steps 20 --warmup 5 --no-extras --shard guides >> $out 2>> $err     # strong, guides sharded (NCCL all-gather of totals)
if [ "$n" = "8" ]; then
  tr --steps 20 --warmup 5 --workload fused --shard database >> $out 2>> $err   # configs[4]
  python bench.py --single-process --gpus $n --steps 20 --warmup 5 >> $out 2>> $err
  tr --steps 20 --warmup 5 --no-extras --scaling weak >> $out 2>> $err     # 100 000 guides per GPU
  nvidia-smi topo -m > gpurun_out/r3_topo.txt 2>&1
fi
wc -l $out
python - <<P
import json
for l in open("$out"):
    d=json.loads(l)
    print(d["n_gpus"], d["config"].get("parallelism", d["config"].get("shard_mode",""))[:40], "value %.4g ms %.3f | e2e %.4g ms %.3f" % (d["value"], d["ms_per_step"], d.get("e2e",{}).get("value",0), d.get("e2e",{}).get("ms_per_step",0)), d.get("roofline",{}).get("step_breakdown_ms"), d["config"].get("sharded_rows_equal_single_gpu_rows"), d["config"].get("shard_fallback_reason"))
P
