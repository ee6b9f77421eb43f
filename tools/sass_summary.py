"""Per-kernel counts of the Blackwell-specific SASS mnemonics in libadm_b200.so (cuobjdump -sass): UTC*MMA = tcgen05.mma,
LDTM / STTM = tcgen05.ld / st, UTMALDG = TMA tensor loads, UTCBAR = tcgen05.commit.  Usage: python tools/sass_summary.py"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "adm_b200", "libadm_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
keys = ["UTCHMMA", "UTCHMMA.2CTA", "LDTM", "STTM", "UTMALDG", "UTCBAR", "HMMA", "MUFU.EX2", "MUFU.TANH"]
counts = collections.OrderedDict()
cur = None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
        counts[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    for k in keys:
        if re.search(r"\b" + re.escape(k) + r"\b", line):
            counts[cur][k] += 1
print(f"# {os.path.relpath(so, ROOT)}: SASS mnemonic counts per kernel (sm_100a)")
print(f"{'kernel':58s} " + " ".join(f"{k:>12s}" for k in keys))
for name, c in counts.items():
    if any(c[k] for k in keys):
        print(f"{name[:58]:58s} " + " ".join(f"{c[k]:12d}" for k in keys))
