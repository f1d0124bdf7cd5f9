"""Dump one kernel's SASS with execution counts from an ncu report (source page).
usage: ncu_sass.py <report.ncu-rep> <kernel-regex> > out.txt"""
import csv, io, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kern],
                     capture_output=True, text=True).stdout
lines = out.splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
end = next((i for i in range(start + 1, len(lines)) if lines[i].startswith('"Kernel Name"')), len(lines))
rows = list(csv.DictReader(io.StringIO("\n".join(lines[start:end]))))
tot = sum(int(r["Instructions Executed"] or 0) for r in rows)
print(f"# {lines[start-1][:120]}  warp-instructions {tot}")
for i, r in enumerate(rows):
    print(f"{i:5d} {int(r['Instructions Executed'] or 0):10d} {float(r['Avg. Threads Executed'] or 0):5.1f} {int(r['# Samples'] or 0):6d}  {r['Source'].strip()[:100]}")
