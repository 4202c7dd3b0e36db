#!/usr/bin/env bash
# Multi-GPU A/B of the halo transports on the 10M/200M graph (run on an N-GPU box):
#   gpurun --gpus 8 -- 'bash tools/scale_sweep.sh 8'
# Writes one JSON line per variant to gpurun_out/scale_sweep_N<N>.jsonl.
set -u
N=${1:-8}
PORT=29600
OUT=gpurun_out/scale_sweep_N${N}.jsonl
: > "$OUT"
run() {
  PORT=$((PORT + 1))
  python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 --master-port "$PORT" \
    bench.py --gpus "$N" --steps 10 --warmup 3 "$@" 2> gpurun_out/scale_sweep.err | grep '^{' | \
    python -c "import sys,json; d=json.loads(sys.stdin.read()); d['_args']=' '.join(sys.argv[1:]); print(json.dumps(d))" "$@" >> "$OUT"
}
run                                   # default: one pull kernel forward, pushed backward
run --fwd packed                      # owner-side pack + copy-engine fetch, 4 stages
run --fwd packed --fwd-stages 6
run --bwd fetch                       # copy-engine fetch of the owner slices into staging + one reduce
run --fwd packed --bwd fetch          # both transports on copy engines
run --bwd pipeline                    # previous backward (copy-engine pulls per owner slice)
python - "$OUT" <<'PY'
import json, sys
for l in open(sys.argv[1]):
    d = json.loads(l)
    print(f"{d['_args'] or '(default)':32s} {d['ms_per_step']:.3f} ms  {d['value']/1e9:.2f} G edges/s  phases rank0 {d['phases_ms_per_rank']['ranks'][0]}")
PY
