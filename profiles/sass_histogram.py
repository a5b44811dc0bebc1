"""Opcode counts per kernel of the built library (cuobjdump -sass), written to profiles/<out>.  usage: python profiles/sass_histogram.py [out]"""
import collections, os, re, subprocess, sys
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(root, "algames.jl_b200", "libalgames_b200.so")
out = os.path.join(root, "profiles", sys.argv[1] if len(sys.argv) > 1 else "sass_histogram.txt")
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
names = {}
try:
    filt = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", txt)), capture_output=True, text=True).stdout.split("\n")
    names = dict(zip(re.findall(r"Function : (\S+)", txt), filt))
except FileNotFoundError:
    pass
special = ("UBLKCP", "SYNCS", "NANOSLEEP", "UTMALDG", "UTMASTG", "DMMA", "UTCMMA", "LDTM", "STTM", "LDGSTS")
rows, cur, cnt = [], None, None
for line in txt.split("\n"):
    mfn = re.search(r"Function : (\S+)", line)
    if mfn:
        if cur: rows.append((cur, cnt))
        cur, cnt = mfn.group(1), collections.Counter()
        continue
    mop = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
    if mop and cur: cnt[mop.group(1)] += 1
if cur: rows.append((cur, cnt))
with open(out, "w") as f:
    f.write("# cuobjdump -sass algames.jl_b200/libalgames_b200.so — opcode counts per kernel (sm_100a)\n"
            "# DFMA/DMUL/DADD/DSETP = FP64 vector pipe; UBLKCP = cp.async.bulk (TMA engine, 1-D) with SYNCS = mbarrier operations; no DMMA /\n"
            "# UTC*MMA: the solve is FP64 with 12x12..24x24 blocks and compile-time sparsity (DESIGN.md §3), tcgen05 has no f64 kind\n\n")
    tot = collections.Counter()
    for fn, c in rows:
        nm = names.get(fn, fn)
        nm = re.sub(r"\(.*", "", nm)
        top = ", ".join(f"{k} {v}" for k, v in c.most_common(12))
        sp = ", ".join(f"{k} {c[k]}" for k in special if c[k])
        f.write(f"{nm:<66} {sum(c.values()):>6} instr: {top}" + (f" | {sp}" if sp else "") + "\n")
        tot.update(c)
    f.write("\nTOTAL " + ", ".join(f"{k} {v}" for k, v in tot.most_common(30)) + "\n")
    f.write("SPECIAL " + ", ".join(f"{k} {tot[k]}" for k in special) + "\n")
print(out)
