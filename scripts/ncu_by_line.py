"""Per-source-line executed warp instructions of the first kernel in an ncu report (--import-source on, -lineinfo):
python scripts/ncu_by_line.py REPORT.ncu-rep UNITS FILE [MIN]"""
import collections, csv, io, os, subprocess, sys
rep, units, want = sys.argv[1], float(sys.argv[2]), sys.argv[3]
mn = float(sys.argv[4]) if len(sys.argv) > 4 else 1.0
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     stdout=subprocess.PIPE, text=True).stdout
cur, hdr, kern = None, None, None
per_line = collections.Counter()
for r in csv.reader(io.StringIO(txt)):
    if not r:
        continue
    if r[0] == "File Path":
        cur = os.path.basename(r[1]); continue
    if r[0] == "Function Name":
        if kern is None: kern = r[1]
        elif r[1] != kern: cur = None
        continue
    if r[0] == "Line No":
        hdr = r; ie = hdr.index("Instructions Executed"); continue
    if hdr is None or cur is None or r[0] == "": continue
    try: per_line[(cur, int(r[0]))] += int(r[ie])
    except ValueError: pass
root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "sac_rcbf_b200", "csrc")
src = open(os.path.join(root, want)).read().split("\n")
for (f, l), n in sorted(per_line.items()):
    if f == want and n / units >= mn:
        print("%5d %7.1f  %s" % (l, n / units, src[l - 1][:110]))
