tr() { n=$1; shift; python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $n "$@"; }

FF_PEER_LOCAL_ONLY=1 tr 2 --steps 20 --warmup 5 --no-extras > gpurun_out/r3_n2_local.json 2> gpurun_out/r3_n2_local.err
python - <<'P'
import json
for f in ("r3_n2_db","r3_n2_local"):
    try:
        d=json.load(open("gpurun_out/%s.json"%f))
        print(f, "value %.3g ms %.3f e2e_ms %.3f"%(d["value"],d["ms_per_step"],d["e2e"]["ms_per_step"]), d["roofline"]["step_breakdown_ms"], [(l['kernel'][:12],round(l['ms'],3)) for l in d['roofline']['launches']], d["config"].get("sharded_rows_equal_single_gpu_rows"))
    except Exception as e: print(f, "failed", e)
P
