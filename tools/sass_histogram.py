"""Opcode histogram of every kernel in libnesosim_b200.so (cuobjdump -sass), written as a markdown table: the
Blackwell-specific mnemonics a reviewer looks for (bulk TMA copies, L2 prefetch, st.async into distributed shared
memory, mbarrier ops, cluster barriers, cp.async, programmatic dependent launch) next to the fp64 / memory mix.
usage: python tools/sass_histogram.py [lib] > profiles/rNN_sass_opcodes.md   (no GPU needed)"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "nesosim_b200", "libnesosim_b200.so")
txt = subprocess.run(["cuobjdump", "-sass", lib], stdout=subprocess.PIPE, text=True, check=True).stdout
KEY = [("UBLKCP", "cp.async.bulk (TMA bulk copy)"), ("UBLKPF", "cp.async.bulk.prefetch.L2"), ("STAS", "st.async (DSMEM + mbarrier tx)"),
       ("SYNCS", "mbarrier arrive / try_wait"), ("UCGABAR_ARV", "barrier.cluster.arrive"), ("UCGABAR_WAIT", "barrier.cluster.wait"),
       ("LDGSTS", "cp.async (global -> shared)"), ("ACQBULK", "griddepcontrol.wait"), ("PREEXIT", "griddepcontrol.launch_dependents"),
       ("MAPA", "mapa (DSMEM address)"), ("DFMA", "fp64 fma"), ("DADD", "fp64 add"), ("DMUL", "fp64 mul"), ("MUFU", "MUFU (rcp64h seed of IEEE division)"),
       ("LDG", "ld.global"), ("STG", "st.global"), ("LDS", "ld.shared"), ("STS", "st.shared"), ("BAR", "bar.sync / bar.arrive"),
       ("ATOM", "atomics (global)"), ("RED", "reductions (global)"), ("LDC", "constant bank loads (LDC/LDCU)")]
funcs = collections.OrderedDict()
cur = None
for ln in txt.splitlines():
    m = re.search(r"Function : (\S+)", ln)
    if m:
        cur = m.group(1)
        funcs[cur] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", ln)
    if m and cur:
        funcs[cur][m.group(1)] += 1
        funcs[cur]["__total__"] += 1


def demangle(n):
    try:
        return subprocess.run(["c++filt", n], stdout=subprocess.PIPE, text=True).stdout.strip()
    except OSError:
        return n


print("# SASS opcode histogram of `%s` (sm_100a, `cuobjdump -sass`)\n" % os.path.basename(lib))
print("Static instruction counts per kernel.  No `UTC*MMA` / `LDTM` anywhere: the path is an fp64 stencil, not a contraction.\n")
print("| kernel | total | " + " | ".join(k for k, _ in KEY) + " |")
print("|---|---|" + "---|" * len(KEY))
for name, c in funcs.items():
    short = re.sub(r"\(.*", "", demangle(name)).replace("nesosim::", "")
    def count(k):
        return sum(v for op, v in c.items() if op == k or (k in ("LDC", "BAR", "ATOM") and op.startswith(k)) or (k == "SYNCS" and op.startswith("SYNCS")))
    print("| `%s` | %d | " % (short, c["__total__"]) + " | ".join(str(count(k)) for k, _ in KEY) + " |")
print("\nLegend: " + "; ".join("`%s` = %s" % kv for kv in KEY) + ".")
