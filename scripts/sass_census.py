"""Opcode census of the built library: `cuobjdump -sass` per kernel, counting the mnemonics that prove the Blackwell-native
path (B200_PROFILING.md: UTC*MMA = tcgen05.mma, LDTM = tcgen05.ld, UTMALDG / UBLKCP = TMA, UTCBAR = tcgen05.commit,
SYNCS = mbarrier) next to the legacy tensor path (HMMA) that must not appear.  Writes a markdown table to stdout.
usage: python scripts/sass_census.py [lib.so] > profiles/r02_sass_census.md"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "petal-neighbors_b200", "lib", "libpetal_b200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
names = subprocess.run(["cu++filt"], input="\n".join(re.findall(r"Function : (\S+)", out)), capture_output=True, text=True).stdout.split("\n")
WATCH = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTCCP", "UTMALDG", "UBLKCP", "UTCBAR", "SYNCS", "HMMA", "FMNMX3", "FMNMX", "FFMA", "FADD", "FMUL",
         "MUFU", "LDG", "LDS", "STS", "ATOM", "RED", "VOTE", "MATCH", "SHFL", "BAR"]
per = collections.OrderedDict()
cur = None
it = iter(names)
for line in out.split("\n"):
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = next(it)
        per[cur] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and cur:
        op = m.group(1)
        per[cur]["_all"] += 1
        for w in WATCH:
            if op == w or op.startswith(w + "."):
                per[cur][w] += 1
                break


def short(n):
    n = re.sub(r"\(.*", "", n)
    return n.replace("petal::", "").replace("void ", "")


tot = collections.Counter()
for c in per.values():
    tot.update(c)
print("# SASS opcode census of `libpetal_b200.so` (sm_100a)\n")
print(f"`cuobjdump -sass` over {len(per)} kernels, {tot['_all']} instructions.  Totals of the watched mnemonics:\n")
print("| mnemonic | count | what it is |\n|---|---|---|")
what = {"UTCHMMA": "tcgen05.mma kind::f16 (5th-gen tensor cores)", "LDTM": "tcgen05.ld (TMEM -> registers)", "UTCCP": "tcgen05.cp (smem -> TMEM)",
        "UTMALDG": "TMA tensor-map load (cp.async.bulk.tensor)", "UBLKCP": "TMA 1-D bulk copy (cp.async.bulk)", "UTCBAR": "tcgen05.commit -> mbarrier",
        "SYNCS": "mbarrier operations", "HMMA": "legacy mma.sync tensor path (must be 0)", "FMNMX3": "3-input min/max (threshold tree)",
        "MATCH": "match.any (hit hand-over)", "VOTE": "ballots (hit compaction)"}
for w in WATCH:
    if tot[w] or w == "HMMA":
        print(f"| `{w}` | {tot[w]} | {what.get(w, '')} |")
print("\nPer kernel family (summed over template instantiations):\n")
fam = collections.OrderedDict()
for n, c in per.items():
    f = re.sub(r"<.*", "", short(n))
    fam.setdefault(f, [0, collections.Counter()])
    fam[f][0] += 1
    fam[f][1].update(c)
cols = ["UTCHMMA", "LDTM", "UTMALDG", "UBLKCP", "UTCBAR", "SYNCS", "FMNMX3", "HMMA"]
print("| kernel | instantiations | instructions | " + " | ".join(cols) + " |\n|---|---|---|" + "---|" * len(cols))
for f, (k, c) in sorted(fam.items(), key=lambda kv: -kv[1][1]["_all"]):
    print(f"| `{f}` | {k} | {c['_all']} | " + " | ".join(str(c[w]) for w in cols) + " |")
