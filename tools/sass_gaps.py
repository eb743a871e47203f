"""static schedule check: stall-count sum between consecutive DMMAs of a kernel (from cuobjdump SASS)"""
import re, subprocess, sys, collections
obj, pat = sys.argv[1], sys.argv[2]
elf = subprocess.run(["cuobjdump", "-elf", obj], capture_output=True, text=True).stdout
names = sorted(set(re.findall(r"_ZN6nipgpu[A-Za-z0-9_]*" + pat + r"[A-Za-z0-9_]*", elf)))
for name in names:
    if name.startswith("_ZN") and "EEv" in name:
        sass = subprocess.run(["cuobjdump", "-sass", "-fun", name, obj], capture_output=True, text=True).stdout
        lines = sass.split("\n"); ins = []; i = 0
        while i < len(lines):
            m = re.match(r"\s+/\*([0-9a-f]{4})\*/\s+(.*?);\s+/\* (0x[0-9a-f]+) \*/", lines[i])
            if m and i + 1 < len(lines):
                m2 = re.match(r"\s+/\* (0x[0-9a-f]+) \*/", lines[i + 1])
                if m2:
                    hi = int(m2.group(1), 16)
                    ins.append((m.group(1), m.group(2).strip(), (hi >> 41) & 0xf, (hi >> 52) & 0x3f)); i += 2; continue
            i += 1
        d = [k for k, x in enumerate(ins) if "DMMA" in x[1]]
        if not d: continue
        # main-loop sweep: the longest run of DMMAs without a branch in between
        gaps = [sum(x[2] for x in ins[a:b]) for a, b in zip(d[:-1], d[1:])]
        span = ins[d[0]:d[-1] + 1]
        br = [t[:30] for a, t, s, w in span if re.search(r"\b(BRA|CALL|BSSY|BSYNC)\b", t)]
        # loop back-edge: find BRA after last DMMA targeting before first DMMA
        print(name[-60:], "DMMAs", len(d), "span instrs", len(span), "stall sum", sum(x[2] for x in span), "branches in span", len(br))
        print("   gaps:", sorted(collections.Counter(gaps).items()))
