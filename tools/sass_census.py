"""SASS census of libdasr_b200.so per kernel (no GPU needed): counts of the mnemonics that prove tcgen05 / TMEM / TMA
(UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld, UTMALDG = TMA load, UTMAREDG = TMA reduce, UTCBAR = tcgen05.commit, SYNCS =
mbarrier, UCGABAR = cluster barrier).   python tools/sass_census.py > profiles/r02_sass_census.txt"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "depth_aware_endoscopy_sr_b200", "libdasr_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], check=True, capture_output=True, text=True).stdout
MN = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAREDG", "UTCBAR", "UTCCP", "SYNCS", "UCGABAR", "LDG", "STG", "REDG", "LDS", "STS"]
per = collections.OrderedDict()
cur = None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        per[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m:
        op = m.group(1)
        per[cur]["_all"] += 1
        for k in MN:
            if op == k or op.startswith(k + ".") or op.startswith(k + "_"):
                per[cur][k] += 1
dem = subprocess.run(["c++filt"], input="\n".join(per.keys()), capture_output=True, text=True).stdout.splitlines()
print("library: %s (%d bytes)" % (os.path.relpath(lib, ROOT), os.path.getsize(lib)))
tot = collections.Counter()
print("%-110s %7s  %s" % ("kernel", "instrs", "  ".join("%s" % k for k in MN)))
for (k, c), name in zip(per.items(), dem):
    name = re.sub(r"\(CUtensorMap_st.*", "(...)", name).replace("dasr::", "")
    print("%-110s %7d  %s" % (name[:110], c["_all"], "  ".join("%*d" % (len(k), c[k]) for k in MN)))
    tot.update(c)
print("%-110s %7d  %s" % ("TOTAL (%d kernels)" % len(per), tot["_all"], "  ".join("%*d" % (len(k), tot[k]) for k in MN)))
