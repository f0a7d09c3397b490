#!/usr/bin/env python
"""Instructions and MUFU of the one-warp scan kernel's tile loop (largest backward-branch region of the SASS;
it includes the mbarrier retry loop and the frame-walk geometry, so it reads ~70 higher than the executed path
quoted in DESIGN.md 3.2 -- the DIFFERENCES between variants are what bench.py's SCAN_TILE_INSTR uses).
    python -m videomamba_b200.build && python tools/scan_tile_instr.py"""
import re, subprocess, sys
out = subprocess.run(["cuobjdump", "-sass", "videomamba_b200/csrc/_obj/scan_fast.o"], capture_output=True, text=True).stdout
funcs = re.split(r"\n\s*Function : ", out)
for f in funcs[1:]:
    name = f.split("\n", 1)[0]
    m = re.search(r"scan1w_kernelILi24ELb0ELb0ELi(\d)ELb0ELb(\d)E", name)
    if not m: continue
    ins = re.findall(r"/\*([0-9a-f]{4,5})\*/\s+(.*?);", f)
    addr = [(int(a, 16), t.strip()) for a, t in ins]
    best = None
    for a, t in addr:
        mm = re.search(r"BRA\S*\s+.*?(0x[0-9a-f]+)", t)
        if mm:
            tgt = int(mm.group(1), 16)
            if tgt < a and (best is None or a - tgt > best[1] - best[0]):
                best = (tgt, a)
    body = [t for a, t in addr if best[0] <= a <= best[1]]
    print(name[60:140], "exp", m.group(1), "gate", m.group(2), "loop instr", len(body), "mufu", sum("MUFU" in t for t in body))
