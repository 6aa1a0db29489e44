"""Run the operand-fetch model (scripts/sass_regmodel.py) on the hot loop of every kernel in a cuobjdump -sass listing.
The hot loop = the backward branch whose body holds the most MUFU.RSQ64H."""
import re, sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from sass_regmodel import model

def functions(path):
    name, cur = None, []
    for ln in open(path):
        m = re.search(r'Function : (\S+)', ln)
        if m:
            if name: yield name, cur
            name, cur = m.group(1), []
            continue
        m = re.match(r'\s*/\*([0-9a-f]{4,5})\*/\s+(.*?);', ln)
        if m and name:
            cur.append((int(m.group(1), 16), m.group(2).strip()))
    if name: yield name, cur

verbose = "-v" in sys.argv
only = [a for a in sys.argv[2:] if not a.startswith("-")]
for name, ins in functions(sys.argv[1]):
    if only and not any(o in name for o in only): continue
    best = None
    for addr, txt in ins:
        m = re.search(r'BRA(?:\.U)?\s+(?:!?U?P\d+,\s*)?0x([0-9a-f]+)', txt)
        if m:
            tgt = int(m.group(1), 16)
            if tgt < addr:
                body = [x for x in ins if tgt <= x[0] <= addr]
                n = sum(1 for x in body if "MUFU.RSQ64H" in x[1])
                if n and (best is None or n > best[0]): best = (n, body)
    if not best:
        print(name, "no loop"); continue
    n, body = best
    tot, n_dp, hist = model(body + body[:0], verbose=verbose)
    print(f"{name}: pairs/iter {n}, DP {n_dp} ({n_dp/n:.1f}/pair), other {len(body)-n_dp}, model cycles {tot} ({tot/n:.1f}/pair), bound {2*n_dp/tot:.4f} -> vs 32-DP peak {64*n/tot:.4f}, fresh hist {hist}")
