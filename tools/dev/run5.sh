timeout 600 python -m pytest tests/test_gpu_multi.py -x -q 2>&1 | tail -5
tr() { n=$1; shift; python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $n "$@"; }
timeout 600 bash -c "$(declare -f tr); tr 2 --steps 20 --warmup 5 --no-extras" > gpurun_out/r3_n2_db.json 2> gpurun_out/r3_n2_db.err
timeout 600 bash -c "$(declare -f tr); tr 2 --steps 20 --warmup 5 --no-extras --shard guides" > gpurun_out/r3_n2_guides.json 2> gpurun_out/r3_n2_guides.err
tail -c 1500 gpurun_out/r3_n2_db.err
python - <<'P'
import json
for f in ("r3_n2_db","r3_n2_guides"):
    try:
        d=json.load(open("gpurun_out/%s.json"%f))
        print(f, "value %.3g ms %.3f e2e %.3g e2e_ms %.3f"%(d["value"],d["ms_per_step"],d["e2e"]["value"],d["e2e"]["ms_per_step"]), d["roofline"]["step_breakdown_ms"], d["config"].get("sharded_rows_equal_single_gpu_rows"), d["config"].get("shard_fallback_reason"))
    except Exception as e: print(f, "failed", e)
P
