"""Summarise '[t4 wait timeout]' lines: distinct (block, site, aux, parity) in order of first appearance, with thread ranges."""
import re, sys, collections
seen = collections.OrderedDict()
for line in sys.stdin:
    m = re.match(r"\[t4 wait timeout\] block (\d+) tid (\d+) site (\d+) aux (-?\d+) parity (\d+)", line)
    if not m:
        if "variant" in line or "rror" in line: print(line.rstrip())
        continue
    b, t, s, a, p = map(int, m.groups())
    seen.setdefault((b, s, a, p), []).append(t)
for (b, s, a, p), t in list(seen.items())[:40]:
    print(f"block {b} site {s} aux {a} parity {p}: {len(t)} threads, tid {min(t)}..{max(t)}")
