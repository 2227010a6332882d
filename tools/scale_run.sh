#!/bin/bash
# The round's multi-GPU measurements (run with: gpurun --gpus 8 -- bash tools/scale_run.sh).  One JSON line per run.
out=gpurun_out/r2_scale.jsonl
: > $out
tr() { n=$1; shift; if [ $n -eq 1 ]; then python bench.py --gpus 1 "$@"; else python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $n "$@"; fi; }
for n in 1 2 4 8; do tr $n --steps 20 --warmup 5 --no-extras >> $out 2>> gpurun_out/r2_scale.err; done
for n in 2 8; do tr $n --steps 20 --warmup 5 --no-extras --scaling weak >> $out 2>> gpurun_out/r2_scale.err; done
tr 8 --steps 20 --warmup 5 --workload fused >> $out 2>> gpurun_out/r2_scale.err
tr 8 --steps 5 --warmup 3 --workload bulge >> $out 2>> gpurun_out/r2_scale.err
tr 1 --steps 20 --warmup 5 --workload fused >> $out 2>> gpurun_out/r2_scale.err
for n in 1 2 4 8; do python bench.py --single-process --gpus $n --steps 20 --warmup 5 >> $out 2>> gpurun_out/r2_scale.err; done
nvidia-smi topo -m > gpurun_out/r2_topo.txt 2>&1
wc -l $out
