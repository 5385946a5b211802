#!/bin/bash
# usage: tools/run_multi.sh NGPUS [bench args...]   (runs bench.py under torchrun, prints a short summary)
N=$1; shift
OUT=gpurun_out/bench_multi_${N}_$$.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 300)) bench.py --gpus $N "$@" > $OUT 2> $OUT.err
echo "exit=$?"
grep -v "OMP_NUM_THREADS\|^\*\*\*\*" $OUT.err | tail -25
python - "$OUT" <<'PY'
import json, sys
try:
    d = json.loads([l for l in open(sys.argv[1]).read().splitlines() if l.startswith("{")][-1])
    print("n_gpus", d["n_gpus"], "ms/step", round(d["ms_per_step"], 4), "value", round(d["value"] / 1e9, 4), "G/s", d["config"]["parallelism"])
    print({k: round(v, 4) for k, v in d["stats"]["stage_ms"].items()})
    print({k: round(v, 4) for k, v in d["stats"]["phase_ms"].items()})
    print("e2e ms", round(d["e2e"]["ms_per_step"], 4))
except Exception as e:
    print("no json:", e)
PY
