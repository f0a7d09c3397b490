#!/usr/bin/env python
"""Pretty-print the key fields of bench.py JSON lines read from stdin or files."""
import json
import sys

srcs = [open(p) for p in sys.argv[1:]] or [sys.stdin]
for f in srcs:
    for line in f:
        line = line.strip()
        if not line.startswith("{"):
            continue
        d = json.loads(line)
        st = {k: round(v["ms_per_step"], 2) for k, v in (d.get("stages") or {}).items()}
        r = d.get("roofline") or {}
        print(f"value {d['value']:.1f} {d['unit']}  ms/step {d['ms_per_step']:.2f}  serial "
              f"{(d.get('config') or {}).get('serial_ms_per_step')}  e2e {d['e2e']['value']:.1f}  "
              f"launches {d.get('gpu_launches')}  scan frac {r.get('frac')}  stages {st}")
