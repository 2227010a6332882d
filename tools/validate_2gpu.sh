python -m pytest tests/test_gpu_multi.py tests/test_gpu_jni.py -x -q 2>&1 | tail -3
python bench.py --single-process --gpus 2 --steps 20 --warmup 5 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('single-process db n2', d['value'], d['ms_per_step'], d['config']['shard_mode'])"
python bench.py --single-process --gpus 2 --steps 20 --warmup 5 --shard guides 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('single-process guides n2', d['value'], d['ms_per_step'], d['config']['shard_mode'])"
