#!/bin/bash
# BASELINE config E on N GPUs: bench.py (strong scaling, parity gate included) for several query-batch sizes.
#   tools/sweep_multi.sh N OUT.jsonl [ROWS]
N=$1; OUT=$2; ROWS=${3:-100000000}
: > "$OUT"
for NQ in 1 16 256 1024 4096 16384; do
  if [ "$N" = "1" ]; then
    timeout 600 python bench.py --gpus 1 --steps 5 --warmup 3 --rows $ROWS --nq $NQ --no-extras --no-cpu-baseline >> "$OUT" 2>> "$OUT.err"
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600 + NQ % 97)) \
      bench.py --gpus $N --steps 5 --warmup 3 --rows $ROWS --nq $NQ --no-extras --no-cpu-baseline >> "$OUT" 2>> "$OUT.err"
  fi
done
python - "$OUT" <<'PY'
import json, sys
for ln in open(sys.argv[1]):
    ln = ln.strip()
    if not ln.startswith("{"):
        continue
    b = json.loads(ln)
    r = b["roofline"]
    print(f"N={b['n_gpus']} rows={b['config']['rows']} nq={b['config']['nq']}: {b['ms_per_step']:.3f} ms/step, {b['value']:.0f} q/s "
          f"(e2e {b['e2e']['value']:.0f}), {r['kernel']} {r['achieved']:.0f} {r['unit']} = {100 * r['frac']:.0f}% of {r['bound']} peak, parity {b['parity']['status']}")
PY
