"""Per-kernel table of the Blackwell-specific SASS opcodes in libsamplenet_b200.so (cuobjdump -sass), written to profiles/.
Runs without a GPU.   python tools/sass_table.py > profiles/r2_sass_opcodes.txt"""
import collections, os, re, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "samplenet_b200", "lib", "libsamplenet_b200.so")
KEYS = ["UTCHMMA", "UTCQMMA", "UTCBAR", "LDTM", "STTM", "UTCCP", "UBLKCP", "UTMALDG", "SYNCS", "ACQBULK", "PREEXIT", "LDGSTS", "HMMA", "FFMA", "MUFU.EX2", "DFMA",
        "ATOMG", "REDG", "RED.", "BAR.SYNC", "SHFL", "LDS", "STS", "LDG", "STG", "LDL", "STL"]
out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
kern = None
tab = collections.OrderedDict()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        kern = re.sub(r"\(.*", "", kern)
        tab[kern] = collections.Counter()
        continue
    if kern is None:
        continue
    m = re.search(r"/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if not m:
        continue
    op = m.group(1)
    tab[kern]["_total"] += 1
    for k in KEYS:
        if op.startswith(k) or (k.endswith(".") and op.startswith(k)):
            tab[kern][k] += 1
cols = [k for k in KEYS if any(t[k] for t in tab.values())]
print("SASS opcode counts per kernel (cuobjdump -sass samplenet_b200/lib/libsamplenet_b200.so; static instruction counts, sm_100a)")
print("tcgen05.mma -> UTCHMMA (kind::tf32/f16), tcgen05.ld/st -> LDTM/STTM, tcgen05.commit -> UTCBAR, cp.async.bulk -> UBLKCP, mbarrier -> SYNCS,")
print("griddepcontrol -> ACQBULK/PREEXIT; LDL/STL = local-memory (spill) traffic\n")
w = max(len(k) for k in tab) + 1
print("%-*s %7s " % (w, "kernel", "instrs") + " ".join("%8s" % c for c in cols))
for k, t in tab.items():
    print("%-*s %7d " % (w, k, t["_total"]) + " ".join("%8s" % (t[c] if t[c] else ".") for c in cols))
