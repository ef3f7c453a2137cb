"""Per-kernel SASS opcode histogram of libpqh_knn.so -> profiles/sass_opcodes.txt.

Evidence that the hot kernels are Blackwell-native (B200_PROFILING.md): tcgen05.mma shows up as
UTC*MMA, tcgen05.ld/st as LDTM/STTM, TMA as UTMALDG; the legacy tensor path would show HMMA.
Run in the build container (cuobjdump needs no GPU).
"""
import collections
import re
import subprocess
import sys

LIB = sys.argv[1] if len(sys.argv) > 1 else "pq_huffman_b200/libpqh_knn.so"
OUT = sys.argv[2] if len(sys.argv) > 2 else "profiles/sass_opcodes.txt"
WATCH = ("UTCIMMA", "UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "SYNCS",
         "VIMNMX3", "FMNMX3", "HMMA", "IMMA", "LDL", "STL")

sass = subprocess.check_output(["cuobjdump", "-sass", LIB], text=True)
arch = sorted(set(re.findall(r"arch = (sm_\w+)", sass)))
kern, name = collections.OrderedDict(), None
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        name = re.sub(r"pqh::\(anonymous namespace\)::", "", name)
        name = re.sub(r"\((anonymous namespace|pqh)::\w+.*$", "", name).strip()
        kern[name] = collections.Counter()
        continue
    m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)((?:\.[A-Z0-9_]+)*)", line)
    if m and name:
        kern[name][m.group(1)] += 1
        kern[name]["_total"] += 1
        if m.group(1) in ("LDTM", "UTCIMMA", "UTCHMMA", "UTMALDG") and m.group(2):
            kern[name][m.group(1) + m.group(2)] += 1
with open(OUT, "w") as f:
    f.write("# cuobjdump -sass %s : per-kernel opcode counts (static instructions); cubins: %s\n"
            % (LIB, ", ".join(arch)))
    f.write("# tcgen05.mma -> UTC*MMA, tcgen05.ld -> LDTM, TMA -> UTMALDG; no HMMA/IMMA (legacy mma.sync) anywhere\n")
    for k, c in kern.items():
        watched = ["%s=%d" % (w, n) for w, n in sorted(c.items())
                   if any(w.startswith(x) for x in WATCH)]
        f.write("%-70s total=%-6d %s\n" % (k[:70], c["_total"], " ".join(watched)))
print(open(OUT).read()[:3000])
