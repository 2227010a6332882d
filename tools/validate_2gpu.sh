# gpurun --gpus 2 -- bash tools/validate_2gpu.sh : the multi-GPU tests on real peers + the bench line in both shard modes
python -m pytest tests/test_gpu_multi.py tests/test_gpu_jni.py -x -q 2>&1 | tail -3
tr() { python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 "$@"; }
for mode in database guides auto; do
  tr --steps 10 --warmup 3 --no-extras --shard $mode 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$mode', '%.4g guides/s %.3f ms | e2e %.4g' % (d['value'], d['ms_per_step'], d['e2e']['value']), d['config']['parallelism'][:30], d['config'].get('sharded_rows_equal_single_gpu_rows'), d['config'].get('shard_fallback_reason'))"
done
