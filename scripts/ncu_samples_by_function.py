"""Warp-stall samples and executed instructions per source function (first kernel of an ncu report captured with
--import-source on): python scripts/ncu_samples_by_function.py REPORT.ncu-rep"""
import bisect, collections, csv, io, os, re, subprocess, sys
rep = sys.argv[1]
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     stdout=subprocess.PIPE, text=True).stdout
root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "sac_rcbf_b200", "csrc")
cur, hdr, kern = None, None, None
ins, smp = collections.Counter(), collections.Counter()
for r in csv.reader(io.StringIO(txt)):
    if not r: continue
    if r[0] == "File Path":
        cur = os.path.basename(r[1]); continue
    if r[0] == "Function Name":
        if kern is None: kern = r[1]
        elif r[1] != kern: cur = None
        continue
    if r[0] == "Line No":
        hdr = r; ie = hdr.index("Instructions Executed"); isamp = hdr.index("# Samples"); continue
    if hdr is None or cur is None or r[0] == "": continue
    try:
        ins[(cur, int(r[0]))] += int(r[ie]); smp[(cur, int(r[0]))] += int(r[isamp])
    except ValueError: pass
pat = re.compile(r'^\s*(?:template.*>\s*)?(?:static\s+)?(?:RCBF_HDC?|__device__(?: __forceinline__)?|__global__)[^;]*?\b([A-Za-z_]\w*)\s*\(')
funcs = {}
for f in {f for f, _ in ins}:
    p = os.path.join(root, f)
    if os.path.exists(p):
        st = []
        for i, l in enumerate(open(p), 1):
            m = pat.match(re.sub(r'__launch_bounds__\([^)]*\)', '', l)) or re.match(r'^(k_\w+)\(', l)
            if m and not l.rstrip().endswith(';'): st.append((i, m.group(1)))
        funcs[f] = st
ai, as_ = collections.Counter(), collections.Counter()
for (f, l), n in ins.items():
    st = funcs.get(f); name = "-"
    if st:
        k = bisect.bisect_right([s[0] for s in st], l) - 1
        name = st[k][1] if k >= 0 else "?"
    ai[(f, name)] += n; as_[(f, name)] += smp[(f, l)]
ti, ts = sum(ai.values()), sum(as_.values())
print("kernel:", kern)
print("%-26s %-28s %7s %7s  %s" % ("file", "function", "inst%", "samp%", "samples per instruction (relative)"))
for k, n in ai.most_common(30):
    print("%-26s %-28s %6.1f%% %6.1f%%  %.2f" % (k[0], k[1], 100.0 * n / ti, 100.0 * as_[k] / ts, (as_[k] / ts) / (n / ti)))
