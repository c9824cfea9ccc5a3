#!/usr/bin/env python
"""Static SASS mnemonic histogram of selected kernels of libpmstep.so (cuobjdump -sass; no GPU needed).
    python scratch/sass_mix.py OUT.txt REGEX [REGEX ...]
Proves which hardware paths a kernel uses: UBLKCP = cp.async.bulk (the TMA engine's 1-D copy), SYNCS =
mbarrier operations, LDGSTS = cp.async, ATOMS = shared-memory atomics, REDG/ATOMG = global atomics."""
import collections, re, subprocess, sys, os
here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(here, "cosmological_particle_mesh_simulation_b200", "libpmstep.so")
out, pats = sys.argv[1], [re.compile(p) for p in sys.argv[2:]]
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
fn, mix, full = None, collections.defaultdict(collections.Counter), collections.defaultdict(collections.Counter)
for line in txt.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        fn = m.group(1) if any(p.search(m.group(1)) for p in pats) else None
        continue
    if fn is None:
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m:
        op = m.group(1)
        mix[fn][op.split(".")[0]] += 1
        full[fn][op] += 1
with open(out, "w") as f:
    for k in mix:
        name = subprocess.run(["c++filt", k], capture_output=True, text=True).stdout.strip() or k
        tot = sum(mix[k].values())
        f.write(f"== {name}\n   {tot} static instructions\n")
        f.write("   " + "  ".join(f"{op} {c}" for op, c in mix[k].most_common(40)) + "\n")
        special = {op: c for op, c in full[k].items() if re.match(r"(UBLKCP|UTMA|SYNCS|LDGSTS|ATOMS|ATOMG|REDG|RED|MATCH|BAR|DEPBAR|ARRIVES|CCTL|FENCE|MEMBAR|ERRBAR|LDS\.128|LDG\.E\.128|STG\.E\.128|UCGABAR)", op)}
        f.write("   async / atomic / barrier forms: " + "  ".join(f"{op} {c}" for op, c in sorted(special.items())) + "\n\n")
print(open(out).read())
