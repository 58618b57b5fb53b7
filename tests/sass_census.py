"""SASS census of libhifigan_b200.so: per kernel, the counts of the mnemonics that prove a Blackwell-native path
(B200_PROFILING.md "What proves a Blackwell-native kernel"): UTC*MMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st,
UTMALDG = TMA tensor loads, UTCBAR = tcgen05.commit, LDGSTS = cp.async, HMMA = legacy mma.sync (must be 0).
Runs on the CPU-only build box:  python tests/sass_census.py > profiles/r02_sass_census.txt"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "hifi-gan_b200", "libhifigan_b200.so")
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
KEYS = ("UTCHMMA", "LDTM", "STTM", "UTMALDG", "UTCBAR", "LDGSTS", "HMMA", "SYNCS", "UCGABAR")
cur, counts = None, collections.OrderedDict()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        name = re.sub(r"\(anonymous namespace\)::", "", name).split("(")[0].replace("void ", "")
        cur = counts.setdefault(name, collections.Counter())
        continue
    if cur is None:
        continue
    m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if not m:
        continue
    op = m.group(1)
    cur["_total"] += 1
    for k in KEYS:
        if op == k or op.startswith(k + "."):
            cur[k] += 1
    if op.startswith("UTCHMMA") and ".2CTA" in op:
        cur["UTCHMMA.2CTA"] += 1
print(f"# {os.path.relpath(so, ROOT)}: {len(counts)} kernels")
print("%-58s %7s %8s %6s %5s %5s %8s %7s %7s %5s" % ("kernel", "instrs", "UTCHMMA", ".2CTA", "LDTM", "STTM", "UTMALDG", "UTCBAR", "LDGSTS", "HMMA"))
tot = collections.Counter()
for name, c in counts.items():
    print("%-58s %7d %8d %6d %5d %5d %8d %7d %7d %5d" % (name[:58], c["_total"], c["UTCHMMA"], c["UTCHMMA.2CTA"], c["LDTM"], c["STTM"],
                                                          c["UTMALDG"], c["UTCBAR"], c["LDGSTS"], c["HMMA"]))
    tot.update(c)
print("%-58s %7d %8d %6d %5d %5d %8d %7d %7d %5d" % ("TOTAL", tot["_total"], tot["UTCHMMA"], tot["UTCHMMA.2CTA"], tot["LDTM"], tot["STTM"],
                                                      tot["UTMALDG"], tot["UTCBAR"], tot["LDGSTS"], tot["HMMA"]))
