"""Attribute a kernel's executed warp instructions to source functions from an ncu report captured with
--import-source on (kernels compiled with -lineinfo):  python scripts/ncu_by_function.py REPORT.ncu-rep UNITS
UNITS = number of work units (e.g. 32-instance tiles) to normalise by."""
import bisect, collections, csv, io, os, re, subprocess, sys

rep, units = sys.argv[1], float(sys.argv[2])
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     stdout=subprocess.PIPE, text=True).stdout
root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "sac_rcbf_b200", "csrc")
cur, hdr, kern = None, None, None
per_line = collections.Counter()
for r in csv.reader(io.StringIO(txt)):
    if not r:
        continue
    if r[0] == "File Path":
        cur = os.path.basename(r[1]); continue
    if r[0] == "Function Name":
        if kern is None:
            kern = r[1]
        elif r[1] != kern:
            cur = None      # only the first kernel in the report
        continue
    if r[0] == "Line No":
        hdr = r; ie = hdr.index("Instructions Executed"); continue
    if hdr is None or cur is None or r[0] == "":
        continue
    try:
        per_line[(cur, int(r[0]))] += int(r[ie])
    except ValueError:
        pass
funcs = {}
pat = re.compile(r'^\s*(?:template.*>\s*)?(?:static\s+)?(?:RCBF_HDC?|__device__(?: __forceinline__)?|__global__)[^;]*?\b([A-Za-z_]\w*)\s*\(')
for f in {f for f, _ in per_line}:
    p = os.path.join(root, f)
    if os.path.exists(p):
        st = []
        for i, l in enumerate(open(p), 1):
            m = pat.match(re.sub(r'__launch_bounds__\([^)]*\)', '', l)) or re.match(r'^(k_\w+)\(', l)
            if m and not l.rstrip().endswith(';'):
                st.append((i, m.group(1)))
        funcs[f] = st
agg = collections.Counter()
for (f, l), n in per_line.items():
    st = funcs.get(f)
    name = "-"
    if st:
        k = bisect.bisect_right([s[0] for s in st], l) - 1
        name = st[k][1] if k >= 0 else "?"
    agg[(f, name)] += n
tot = sum(agg.values())
print("kernel:", kern)
print("warp instructions executed: %d = %.1f per unit (%g units)" % (tot, tot / units, units))
for (f, name), n in agg.most_common(40):
    print("%-26s %-28s %8.1f  %5.1f%%" % (f, name, n / units, 100.0 * n / tot))
