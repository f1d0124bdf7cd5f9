"""Top stall locations of one kernel from an ncu report's SASS source page.
usage: ncu_hot.py <report.ncu-rep> <kernel-regex> [top-n]"""
import csv, subprocess, sys, io
rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kern],
                     capture_output=True, text=True).stdout
lines = out.splitlines()
# first kernel only
start = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
end = next((i for i in range(start + 1, len(lines)) if lines[i].startswith('"Kernel Name"')), len(lines))
rows = list(csv.DictReader(io.StringIO("\n".join(lines[start:end]))))
tot = sum(int(r["# Samples"] or 0) for r in rows)
print(f"{lines[start-1][:100]}  total samples {tot}, instructions {len(rows)}")
stalls = [k for k in rows[0] if k.startswith("stall_") and "Not Issued" not in k]
agg = {k: sum(int(r[k] or 0) for r in rows) for k in stalls}
print("stall mix:", ", ".join(f"{k[6:]} {100*v/tot:.1f}%" for k, v in sorted(agg.items(), key=lambda x: -x[1]) if v * 50 > tot))
idx = sorted(range(len(rows)), key=lambda i: -int(rows[i]["# Samples"] or 0))[:top]
for i in sorted(idx):
    r = rows[i]
    s = int(r["# Samples"] or 0)
    main = max(stalls, key=lambda k: int(r[k] or 0))
    print(f"{i:5d} {100*s/tot:5.1f}%  {main[6:]:12s} wf={r['L1 Wavefronts Shared']:>9s} ideal={r['L1 Wavefronts Shared Ideal']:>9s}  {r['Source'].strip()[:90]}")
if len(sys.argv) > 4:
    bounds = [int(x) for x in sys.argv[4].split(",")]
    bounds = [0] + bounds + [len(rows)]
    for a, b in zip(bounds[:-1], bounds[1:]):
        seg = rows[a:b]
        s = sum(int(r["# Samples"] or 0) for r in seg)
        ins = sum(int(r["Instructions Executed"] or 0) for r in seg)
        wf = sum(int(r["L1 Wavefronts Shared"] or 0) for r in seg)
        wfi = sum(int(r["L1 Wavefronts Shared Ideal"] or 0) for r in seg)
        mix = {k: sum(int(r[k] or 0) for r in seg) for k in stalls}
        ms = ", ".join(f"{k[6:]} {100*v/max(s,1):.0f}%" for k, v in sorted(mix.items(), key=lambda x: -x[1])[:4])
        print(f"[{a:5d},{b:5d}) samples {100*s/tot:5.1f}%  warp-instr {ins/1e6:8.1f}M  smem wf {wf/1e6:7.1f}M (ideal {wfi/1e6:7.1f}M)  {ms}")
