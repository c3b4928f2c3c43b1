"""Per-line warp-stall samples (+ executed instructions) of one source file for the first kernel in an ncu report:
python scripts/ncu_samples_by_line.py REPORT.ncu-rep FILE [MIN_PCT]"""
import collections, csv, io, os, subprocess, sys
rep, want = sys.argv[1], sys.argv[2]
mn = float(sys.argv[3]) if len(sys.argv) > 3 else 0.3
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     stdout=subprocess.PIPE, text=True).stdout
cur, hdr, kern = None, None, None
ins, smp = collections.Counter(), collections.Counter()
stall_cols = {}
stalls = collections.defaultdict(collections.Counter)
for r in csv.reader(io.StringIO(txt)):
    if not r: continue
    if r[0] == "File Path":
        cur = os.path.basename(r[1]); continue
    if r[0] == "Function Name":
        if kern is None: kern = r[1]
        elif r[1] != kern: cur = None
        continue
    if r[0] == "Line No":
        hdr = r; ie = hdr.index("Instructions Executed"); isamp = hdr.index("# Samples")
        stall_cols = {i: h for i, h in enumerate(hdr) if h.startswith("stall_")}
        continue
    if hdr is None or cur is None or r[0] == "": continue
    try:
        key = (cur, int(r[0]))
        ins[key] += int(r[ie]); smp[key] += int(r[isamp])
        for i, h in stall_cols.items():
            if i < len(r) and r[i] not in ("", "0"):
                stalls[key][h[6:]] += int(r[i])
    except ValueError: pass
root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "sac_rcbf_b200", "csrc")
src = open(os.path.join(root, want)).read().split("\n")
ts = sum(smp.values())
for (f, l), n in sorted(smp.items()):
    if f == want and 100.0 * n / ts >= mn:
        top = ",".join("%s:%d" % kv for kv in stalls[(f, l)].most_common(3))
        print("%5d %5.2f%% %-44s %s" % (l, 100.0 * n / ts, top, src[l - 1].strip()[:80]))
