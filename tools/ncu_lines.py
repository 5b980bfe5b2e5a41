"""Samples per CUDA source line of one kernel: ncu -i rep --page source --csv --print-source cuda,sass | python tools/ncu_lines.py [top]"""
import csv, sys
top = int(sys.argv[1]) if len(sys.argv) > 1 else 50
cur_file, hdr, lines = None, None, []
for row in csv.reader(sys.stdin):
    if not row: continue
    if row[0] == "File Path": cur_file = row[1].split("/")[-1]; continue
    if row[0] == "Function Name": continue
    if row[0] == "Line No": hdr = row; continue
    if hdr and row[0]:
        # source text may hold unescaped quotes / commas: index the numeric columns from the END of the row
        si = hdr.index("# Samples") - len(hdr)
        ei = hdr.index("Instructions Executed") - len(hdr)
        try:
            lines.append((cur_file, int(row[0]), row[1].strip(), int(row[si] or 0), int(row[ei] or 0)))
        except ValueError:
            pass
tot = sum(l[3] for l in lines)
print("total samples", tot)
for f, n, src, s, e in sorted(lines, key=lambda l: -l[3])[:top]:
    print(f"{100 * s / tot:5.1f}%  exec {e:>9d}  {f}:{n:<4d} {src[:100]}")
