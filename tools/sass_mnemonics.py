"""SASS mnemonic counts per kernel of guassianhand_b200/libghr.so (cuobjdump -sass), restricted to the mnemonics
that show which hardware paths the kernels use.   python tools/sass_mnemonics.py > profiles/r2_sass_mnemonics.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "guassianhand_b200", "libghr.so")
KEEP = re.compile(r"^(UBLKCP|SYNCS|FFMA2?$|FMUL2|FADD2|REDG|RED\.|ATOMG|ATOMS|CREDUX|MATCH|LDG\.E\.128|STG\.E\.128|LDS\.128|"
                  r"MUFU|MEMBAR|UTMALDG|UTC|FENCE|SHFL|VOTE)")
out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
arch = sorted(set(re.findall(r"arch = (sm_\w+)", out)))
kern, counts = None, collections.OrderedDict()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        name = re.sub(r"\(anonymous namespace\)::|ghr::|void ", "", name)
        name = re.sub(r"\(.*", "", name)
        kern = name
        counts.setdefault(kern, collections.Counter())
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and kern:
        op = m.group(1)
        if KEEP.match(op):
            counts[kern][op] += 1
print(f"# SASS mnemonic counts per kernel of the shipped guassianhand_b200/libghr.so (cuobjdump -sass; cubins: {', '.join(arch)})")
print(f"# {'kernel':44s} {'mnemonic':36s} count")
for k, c in counts.items():
    for op, n in sorted(c.items()):
        print(f"{k:46s} {op:36s} {n}")
