#!/bin/bash
# Multi-GPU measurement pass (run under `gpurun --gpus N`): NCCL parity test, bench.py weak + strong scaling of
# configs[1], strong scaling of configs[4], and the N-GPU rows of configs[2] / configs[3] (tools/nbench.py).
# One JSON line per run lands in gpurun_out/multi_<tag>_n<N>.jsonl.
set -u
N=${1:-2}
TAG=${2:-r2}
mkdir -p gpurun_out
OUT=gpurun_out/multi_${TAG}_n${N}.jsonl
: > $OUT
run() {  # port, script + args
    local port=$1; shift
    if [ "$N" -eq 1 ]; then timeout 600 python "$@"; else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port "$@"; fi
}
if [ "$N" -ge 2 ]; then
    timeout 900 python -m pytest tests/test_gpu_multirank.py -q -m gpu 2>&1 | tail -3
fi
run 29611 bench.py --gpus $N --steps 2 --warmup 3 --no-cpu-baseline 2> gpurun_out/multi_${TAG}_n${N}_weak.err | grep '^{' | tail -1 >> $OUT
run 29612 bench.py --gpus $N --steps 2 --warmup 3 --no-cpu-baseline --scaling strong 2> gpurun_out/multi_${TAG}_n${N}_strong.err | grep '^{' | tail -1 >> $OUT
run 29613 bench.py --gpus $N --steps 2 --warmup 3 --no-cpu-baseline --scaling strong --config 4 2> gpurun_out/multi_${TAG}_n${N}_c4.err | grep '^{' | tail -1 >> $OUT
run 29614 tools/nbench.py --config 2 --steps 20 2> gpurun_out/multi_${TAG}_n${N}_nb2.err | grep '^{' >> $OUT
run 29615 tools/nbench.py --config 3 --pairs 10000 2> gpurun_out/multi_${TAG}_n${N}_nb3.err | grep '^{' >> $OUT
python - <<EOF
import json
for l in open("$OUT"):
    try: d = json.loads(l)
    except Exception: continue
    if "metric" in d:
        c = d.get("config", {})
        print("bench N=%s %s scaling=%s value=%.1f %s e2e=%.1f ms/step=%.1f equal=%s" % (d["n_gpus"], c.get("workload", "")[:11], d["scaling"], d["value"], d["unit"], d["e2e"]["value"], d["ms_per_step"], c.get("resident_equals_e2e")))
    else:
        print(json.dumps(d)[:300])
EOF
for f in gpurun_out/multi_${TAG}_n${N}_*.err; do tail -n 2 "$f"; done | grep -v "^$" | tail -12
