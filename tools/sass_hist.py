#!/usr/bin/env python3
"""Static SASS opcode histogram of one kernel (substring match on the mangled name) of an object file.
   python tools/sass_hist.py obj name_substr [weights_per_lane_in_loop]"""
import collections, re, subprocess, sys
obj, key = sys.argv[1], sys.argv[2]
wpl = float(sys.argv[3]) if len(sys.argv) > 3 else None
out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
blocks = out.split("Function : ")
for b in blocks[1:]:
    name = b.split("\n", 1)[0]
    if key not in name:
        continue
    ops = collections.Counter()
    for m in re.finditer(r"^\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", b, flags=re.M):
        ops[m.group(1)] += 1
    tot = sum(ops.values())
    print(name[:100], "total", tot)
    print("  ".join(f"{o} {n}" + (f" ({n / wpl:.3f})" if wpl else "") for o, n in ops.most_common(24)))
    break
