#!/usr/bin/env python
"""Per-kernel counts of the SASS mnemonics that show what hardware a kernel uses (tcgen05 MMA / TMEM loads / TMA tensor and
bulk copies / mbarrier transactions / clusters / dp4a), from the built library.  Writes profiles/sass_summary.txt.
    python tools/sass_summary.py"""
import collections, os, re, subprocess, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(REPO, "robot_ebert_b200", "librebert_b200.so")
PAT = {"UTCHMMA (tcgen05.mma f16/bf16)": r"\bUTCHMMA", "UTCIMMA (tcgen05.mma i8)": r"\bUTCIMMA", "UTCQMMA (tcgen05.mma f8f6f4)": r"\bUTCQMMA",
       "2CTA forms (cta_group::2)": r"\.2CTA", "LDTM (tcgen05.ld)": r"\bLDTM", "UTMALDG (TMA tensor load)": r"\bUTMALDG",
       "UBLKCP (TMA bulk copy)": r"\bUBLKCP", "SYNCS (mbarrier)": r"\bSYNCS", "UTCBAR (tcgen05.commit)": r"\bUTCBAR",
       "UCGABAR (cluster barrier)": r"\bUCGABAR", "IDP.4A (dp4a)": r"\bIDP\.4A", "ACQBULK/PDL (griddepcontrol)": r"\bACQBULK|PREEXIT",
       "DFMA (fp64)": r"\bDFMA", "MUFU.RCP64H": r"MUFU\.RCP64H"}
out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
kern, counts = None, collections.OrderedDict()
for ln in out.splitlines():
    m = re.search(r"Function : (\S+)", ln)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        kern = re.sub(r"\(.*", "", kern)
        counts[kern] = collections.Counter()
        continue
    if kern:
        for name, pat in PAT.items():
            if re.search(pat, ln):
                counts[kern][name] += 1
lines = [f"SASS digest of {os.path.relpath(LIB, REPO)} (cuobjdump -sass, sm_100a); regenerate with tools/sass_summary.py", ""]
for k, c in counts.items():
    if not any(c.values()):
        continue
    lines.append(k)
    for name in PAT:
        if c[name]:
            lines.append(f"    {name:34s} {c[name]}")
    lines.append("")
os.makedirs(os.path.join(REPO, "profiles"), exist_ok=True)
with open(os.path.join(REPO, "profiles", "sass_summary.txt"), "w") as fh:
    fh.write("\n".join(lines))
print("\n".join(lines[:60]))
