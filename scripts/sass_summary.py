import subprocess, re, collections
out = subprocess.run(["cuobjdump", "-sass", "pcfa_b200/lib/libpcfa_b200.so"], capture_output=True, text=True).stdout
cols = ["UTCHMMA", "UTCQMMA", "UTMALDG", "UTMASTG", "UTMAREDG", "LDTM", "STTM", "UTCBAR", "UTCCP", "REDG", "LDG", "STG", "LDS", "STS", "FFMA", "HMMA", "REDUX", "ATOMS"]
kern = None
stats = collections.OrderedDict()
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        kern = m.group(1); stats[kern] = collections.Counter(); continue
    m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and kern:
        op = m.group(1)
        stats[kern]["instrs"] += 1
        base = op.split(".")[0]
        if base in cols: stats[kern][base] += 1
        if base == "RED": stats[kern]["REDG"] += 1
dem = subprocess.run(["c++filt"], input="\n".join(stats), capture_output=True, text=True).stdout.splitlines()
rows = []
for name, (k, c) in zip(dem, stats.items()):
    short = re.sub(r"\(.*", "", name)
    rows.append((short, c))
rows.sort(key=lambda r: -r[1]["instrs"])
print("# cuobjdump -sass pcfa_b200/lib/libpcfa_b200.so (sm_100a), instruction-mnemonic counts per kernel (round 2, end state)")
print("# UTCHMMA = tcgen05.mma (.2CTA = cta_group::2), UTMALDG / UTMAREDG = TMA tensor load / reduce-add, LDTM = tcgen05.ld, UTCBAR = tcgen05.commit, REDUX = warp reduce")
print("%-58s %8s " % ("kernel", "instrs") + " ".join("%8s" % c for c in cols))
for short, c in rows:
    print("%-58s %8d " % (short[:58], c["instrs"]) + " ".join("%8d" % c[x] for x in cols))
