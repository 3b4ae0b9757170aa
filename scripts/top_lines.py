"""Per-source-line totals of an ncu report's source page:  ncu -i X.ncu-rep --page source --csv --print-source cuda,sass > x.csv;
python scripts/top_lines.py x.csv [n].  Prints the n source lines with the most executed warp instructions (and their stall samples)."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 30
hdr = None; fname = ""; out = []
for r in rows:
    if r and r[0] in ("File Name", "File Path") and len(r) > 1 and r[1]: fname = r[1].split("/")[-1]; continue
    if "Instructions Executed" in r: hdr = r; continue
    if hdr is None or len(r) < 8 or not r[0].isdigit(): continue
    ie = hdr.index("Instructions Executed"); sm = hdr.index("# Samples")
    try: out.append((int(r[ie]), int(r[sm] or 0), fname, int(r[0])))
    except ValueError: pass
tot = sum(o[0] for o in out); ts = sum(o[1] for o in out)
print("total warp instructions", tot, "samples", ts)
src = {}
for o in sorted(out, reverse=True)[:n]:
    f = "/root/repo/gencore_b200/csrc/" + o[2]
    if f not in src:
        try: src[f] = open(f).read().split("\n")
        except OSError: src[f] = []
    line = src[f][o[3] - 1].strip()[:100] if o[3] - 1 < len(src[f]) else ""
    print("%9d %5.1f%%  samples %5d %5.1f%%  %s:%d  %s" % (o[0], 100.0 * o[0] / tot, o[1], 100.0 * o[1] / max(ts, 1), o[2], o[3], line))
