N=$1
for sc in weak strong; do
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2971$N bench.py --gpus $N --steps 2 --warmup 3 --no-cpu-baseline --scaling $sc 2>/dev/null | grep '^{' | tail -1 >> gpurun_out/multi_r2final_n$N.jsonl
done
python - <<EOF
import json
for l in open("gpurun_out/multi_r2final_n$N.jsonl"):
    d=json.loads(l); print("N=%d %s value=%.1f e2e=%.1f ms=%.1f equal=%s" % (d["n_gpus"], d["scaling"], d["value"], d["e2e"]["value"], d["ms_per_step"], d["config"].get("resident_equals_e2e")))
EOF
