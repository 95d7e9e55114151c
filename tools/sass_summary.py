"""Opcode histogram of the hot kernels of the shipped library (cuobjdump -sass), plus the Blackwell-specific mnemonics.
python tools/sass_summary.py > profiles/r02_sass_summary.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "napa_fft3_b200", "libnapafft_b200.so")
want = sys.argv[1:] or ["fft_pass_kernelILi8ELi0ELb1E", "fft_pass_kernelILi8ELi2ELb1E", "fft_pass_kernelILi12ELi1ELb1E", "fft_pass_kernelILi12ELi1ELb0E",
                        "fft_pass_kernelILi13ELi1ELb0E", "fft_fused_kernelILi8ELi8ELi0ELb0E", "fft_fused_tma_kernelILi8ELi8ELi0ELb0E", "bitrev_tiled_kernelI7double2"]
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
cur, hist = None, {}
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
    if m and cur:
        hist.setdefault(cur, collections.Counter())[m.group(1).split(".")[0]] += 1
special = ("UTMALDG", "UTMASTG", "UBLKCP", "SYNCS", "LDGSTS", "UTCHMMA", "LDTM", "REDG", "MEMBAR", "CCTL", "FENCE", "SHFL", "BAR")
print("# cuobjdump -sass napa_fft3_b200/libnapafft_b200.so -- static opcode counts of the hot kernels (sm_100a)")
for fn, h in hist.items():
    if not any(w in fn for w in want):
        continue
    tot = sum(h.values())
    fp64 = h["DADD"] + h["DMUL"] + h["DFMA"]
    print("\n%s\n  %d instructions, FP64 %d (%.0f %%): DADD %d DMUL %d DFMA %d; LDG %d STG %d LDS %d STS %d" % (
        fn, tot, fp64, 100.0 * fp64 / tot, h["DADD"], h["DMUL"], h["DFMA"], h["LDG"], h["STG"], h["LDS"], h["STS"]))
    print("  other: " + ", ".join("%s %d" % kv for kv in h.most_common(30) if kv[0] not in ("DADD", "DMUL", "DFMA", "LDG", "STG", "LDS", "STS")))
    sp = {k: h[k] for k in special if h[k]}
    print("  async / barrier / fence mnemonics: %s" % (sp or "none"))
