"""Condense the compute-sanitizer logs of tools/sanitize.sh into profiles/sanitizer_<tag>.txt.
usage: sanitize_summary.py <tag> [tools...]   (reads gpurun_out/<tag>_san_<tool>.log and ..._pytest.log)"""
import collections
import os
import re
import sys

tag = sys.argv[1]
tools = sys.argv[2:] or ["memcheck", "racecheck", "synccheck", "initcheck"]
out = [f"# compute-sanitizer over the small GPU parity tests (tools/sanitize.sh {tag}); one section per tool"]
for tool in tools:
    log, pt = f"gpurun_out/{tag}_san_{tool}.log", f"gpurun_out/{tag}_san_{tool}_pytest.log"
    if not os.path.exists(log):
        continue
    text = open(log, errors="replace").read().splitlines()
    summ = [l.strip("= ").strip() for l in text if "SUMMARY" in l]
    kinds = collections.Counter()
    where = collections.Counter()
    cur = None
    for l in text:
        m = re.match(r"=+ (Invalid|Uninitialized|Race|Warning|Error|Potential|Barrier|Program hit)(.*)", l)
        if m:
            cur = (m.group(1) + m.group(2)).strip()[:110]
            kinds[cur] += 1
        m = re.match(r"=+\s+at (\S+)", l)
        if m and cur:
            where[m.group(1)[:90]] += 1
            cur = None
    tests = open(pt, errors="replace").read().strip().splitlines()[-3:] if os.path.exists(pt) else []
    out.append(f"\n[{tool}]")
    out += ["  pytest: " + t for t in tests]
    out += ["  " + s for s in summ[-3:]] or ["  (no summary line: the run was cut off)"]
    for k, n in kinds.most_common(12):
        out.append(f"  {n:6d} x {k}")
    for k, n in where.most_common(12):
        out.append(f"  {n:6d} at {k}")
open(f"profiles/sanitizer_{tag}.txt", "w").write("\n".join(out) + "\n")
print("\n".join(out))
