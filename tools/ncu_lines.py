"""Warp-instructions and stall samples per CUDA source line of one kernel (ncu source page, cuda,sass view).
usage: ncu_lines.py <report.ncu-rep> <kernel-regex> [top-n] > out.txt"""
import csv
import io
import subprocess
import sys

rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 60
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv",
                      "--kernel-name", "regex:" + kern], capture_output=True, text=True).stdout
rows, fname, func, hdr = [], "", "", None


def num(d, k):
    v = (d.get(k) or "0").replace(",", "")
    try:
        return int(float(v))
    except ValueError:
        return 0


for rec in csv.reader(io.StringIO(out)):
    if not rec:
        continue
    if rec[0] == "File Path":
        fname = rec[1].rsplit("/", 1)[-1]
    elif rec[0] == "Function Name":
        func = rec[1]
    elif rec[0] == "Line No":
        hdr = rec
    elif hdr and rec[0] not in ("", "Kernel Name") and rec[0].isdigit():
        d = dict(zip(hdr[2:], rec[-(len(hdr) - 2):]))     # from the end: source lines may hold quotes and commas
        rows.append((fname, int(rec[0]), rec[1].strip(), num(d, "Instructions Executed"), num(d, "# Samples"),
                     num(d, "stall_barrier"), num(d, "stall_long_sb"), num(d, "stall_short_sb"),
                     num(d, "stall_math"), num(d, "L1 Wavefronts Shared"), num(d, "L1 Wavefronts Shared Ideal")))
tot_i = sum(r[3] for r in rows) or 1
tot_s = sum(r[4] for r in rows) or 1
print(f"# {func}: {tot_i} warp-instructions, {tot_s} stall samples; per CUDA source line, top {top} by instructions")
print("#  inst%  samp%  barrier long_sb short_sb  math  smem_wf/ideal  file:line  source")
for r in sorted(rows, key=lambda r: -r[3])[:top]:
    print(f"{100.0 * r[3] / tot_i:6.2f} {100.0 * r[4] / tot_s:6.2f} {r[5]:7d} {r[6]:7d} {r[7]:8d} {r[8]:5d}  {r[9]:9d}/{r[10]:<9d} {r[0]}:{r[1]}  {r[2][:110]}")
