#!/usr/bin/env bash
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out/r2k; mkdir -p "$OUT"
timeout -k 10 120 python bench.py --workload nca --nca-batch 32 --nca-pool 128 --steps 6 --warmup 3 > "$OUT/nca.log" 2> "$OUT/nca.err"; echo "rc=$?"
python - "$OUT/nca.log" <<'PY'
import json, sys
j = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print({k: j.get(k) for k in ("value", "ms_per_step", "warmup", "e2e", "verify", "graph")})
PY
tail -2 "$OUT/nca.err"
