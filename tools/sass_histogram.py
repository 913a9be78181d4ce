"""tools/sass_histogram.py -- per-kernel SASS opcode histogram of the shipped libprealps_cuda.so (cuobjdump -sass), the evidence
for which instructions the hot kernels are made of (DMMA = FP64 tensor core, LDG.E.NA.EFL2.256 = streaming panel loads,
LDGSTS = cp.async, UBLKCP = cp.async.bulk, SYNCS = mbarrier, ACQBULK / PREEXIT = programmatic dependent launch):
    python tools/sass_histogram.py > profiles/r02_sass_histogram.txt"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "prealps_b200", "lib", "libprealps_cuda.so")
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
kernels, cur = collections.OrderedDict(), None
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        name = re.sub(r"\(anonymous namespace\)::", "", name)
        cur = kernels.setdefault(re.sub(r"\(.*$", "", name), collections.Counter())
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
    if m and cur is not None:
        cur[m.group(1)] += 1
KEEP = ("DMMA", "DFMA", "DADD", "LDG", "STG", "LDS", "STS", "LDGSTS", "UBLKCP", "SYNCS", "ACQBULK", "PREEXIT", "ATOM", "RED", "BAR", "MEMBAR", "SHFL", "LDL", "STL")
print("# cuobjdump -sass prealps_b200/lib/libprealps_cuda.so, built with nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo")
print("# instructions per kernel, then the opcodes that matter (all variants of a mnemonic summed; full mnemonics for loads)")
for k, c in kernels.items():
    if not any(x in k for x in ("sweep", "assemble", "spmm", "gram2_mma", "ortho_update_mma", "update_z_mma", "transform_update", "halo")):
        continue
    tot = sum(c.values())
    groups = collections.Counter()
    for op, n in c.items():
        base = op.split(".")[0]
        if base in KEEP:
            groups[op if base in ("LDG", "LDGSTS", "UBLKCP", "DMMA") else base] += n
    print("%s: %d instructions; %s" % (k, tot, ", ".join("%s %d" % kv for kv in sorted(groups.items(), key=lambda kv: -kv[1]))))
