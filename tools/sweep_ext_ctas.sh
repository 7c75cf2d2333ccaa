#!/bin/bash
# resident CTAs per SM for the extension kernel (DN_EXT_CTAS): one bench line per setting
for c in "$@"; do
  DN_EXT_CTAS=$c python bench.py --steps 10 --warmup 3 --cpu-sample-mbp 1 2>/dev/null > /tmp/sweep_$c.json
  python - "$c" <<'PY'
import json, sys
c = sys.argv[1]
d = json.loads(open("/tmp/sweep_%s.json" % c).read().strip().split("\n")[-1])
print(c, round(d["value"], 3), round(d["ms_per_step"], 3), d["stage_ms_per_step"])
PY
done
