"""Per-kernel SASS evidence for the Blackwell-specific instructions (cuobjdump -sass on the built library).
usage: python profiles/sass_summary.py [rpeflow_b200/libb200flow.so] > profiles/r2_sass_summary.md"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "rpeflow_b200", "libb200flow.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
WANT = [("UTCHMMA", "tcgen05.mma (kind::tf32)"), ("LDTM", "tcgen05.ld"), ("UTMALDG", "TMA tensor load"), ("UBLKCP", "cp.async.bulk"),
        ("SYNCS", "mbarrier ops"), ("FFMA2", "packed fp32x2 FMA"), ("FADD2", "packed fp32x2 add"), ("REDUX", "redux.sync"),
        ("LDGSTS", "cp.async"), ("RED", "fire-and-forget atomics"), ("UTCBAR", "tcgen05.commit"), ("ATOMS", "shared-memory atomics")]
kern = None
counts = collections.OrderedDict()
total = collections.Counter()
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        kern = re.sub(r"\(.*", "", name).replace("void ", "")
        counts[kern] = collections.Counter()
        continue
    if kern is None:
        continue
    m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m:
        op = m.group(1)
        counts[kern]["_all"] += 1
        for w, _ in WANT:
            if op == w or (w == "REDUX" and op == "CREDUX"):          # CREDUX = redux.sync min/max (result in a uniform register)
                counts[kern][w] += 1
                total[w] += 1
print("# SASS evidence — `cuobjdump -sass rpeflow_b200/libb200flow.so` (sm_100a), instruction counts per kernel\n")
print("Mnemonics: " + "; ".join(f"`{w}` = {d}" for w, d in WANT) + ".\n")
cols = [w for w, _ in WANT if total[w]]
print("| kernel | SASS instrs | " + " | ".join(cols) + " |")
print("|---|---|" + "---|" * len(cols))
for k, c in counts.items():
    if any(c[w] for w in cols):
        print(f"| `{k[:70]}` | {c['_all']} | " + " | ".join(str(c[w]) if c[w] else "" for w in cols) + " |")
print("\n| total | | " + " | ".join(str(total[w]) for w in cols) + " |")
print(f"\n{len(counts)} kernels in the library; archs: " + subprocess.run(["cuobjdump", "-lelf", lib], capture_output=True, text=True).stdout.strip().replace("\n", ", "))
