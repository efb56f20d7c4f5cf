"""Per-kernel count of the SASS mnemonics that show what a kernel is built on (tcgen05 = UTC*MMA, TMEM loads
= LDTM, TMA = UTMALDG, cluster barriers, shared/global atomics, fp64 FMA): `cuobjdump -sass` of the in-tree
library, summarised.  Usage: python scripts/sass_excerpt.py > profiles/<round>_sass_excerpt.txt"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "autometa_b200", "libautometa_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
pat = re.compile(r'\b(UTCIMMA(?:\.2CTA)?|UTCHMMA|LDTM(?:\.x\d+)?|STTM|UTMALDG(?:\.[\w.]+)?|UTMASTG|UBLKCP|'
                 r'UTCBAR(?:\.[\w.]+)?|ATOMS(?:\.[\w.]+)?|DFMA|DMMA|HMMA|REDG?\.[\w.]+|ATOMG\.[\w.]+|'
                 r'UCGABAR_\w+|LDGSTS(?:\.[\w.]+)?|SHFL\.\w+|MATCH\.\w+)\b')
stats, lines, cur = collections.OrderedDict(), {}, None
for line in sass.splitlines():
    m = re.search(r'Function : (\S+)', line)
    if m:
        cur = m.group(1)
        stats[cur] = collections.Counter()
        lines[cur] = 0
        continue
    if cur and re.search(r'/\*[0-9a-f]{4,}\*/', line):
        lines[cur] += 1
        for t in pat.findall(line):
            stats[cur][t] += 1
print(f"# {os.path.relpath(so, ROOT)}: {len(stats)} kernels (cuobjdump -sass, sm_100a)")
for k, c in stats.items():
    name = subprocess.run(["c++filt", k], capture_output=True, text=True).stdout.strip()
    name = re.sub(r'^void ', '', name)
    name = re.sub(r'\((?:[^()]|\([^()]*\))*\)$', '', name)
    name = name.replace("(anonymous namespace)::", "")
    print(f"{name}  [{lines[k]} instructions]")
    if c:
        print("    " + ", ".join(f"{t} x{n}" for t, n in sorted(c.items())))
