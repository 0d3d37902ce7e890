#!/usr/bin/env python
"""Counts the SASS mnemonics that prove Blackwell-native code paths (tcgen05 / TMEM / TMA) per kernel of the built library.
usage: python tools/sass_grep.py > profiles/rNN_sass_mnemonics.txt   (needs cuobjdump; no GPU)"""
import collections, os, re, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "see-vcn_b200", "csrc", "libseevcn_b200.so")
WANT = ["UTCHMMA", "UTCBAR", "UTMALDG", "UTMASTG", "UTMACMDFLUSH", "UTMAPF", "FENCE.VIEW.ASYNC", "ACQBULK", "LDTM", "STTM", "SYNCS", "REDUX", "MATCH", "ATOMS", "RED.", "ATOMG", "HMMA", "DMMA", "IMMA"]


def main():
    txt = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    funcs = re.split(r"\n\s*Function : ", txt)
    tot, rows = collections.Counter(), []
    for f in funcs[1:]:
        name = f.split("\n", 1)[0].strip()
        c = collections.Counter()
        for m in WANT:
            n = len(re.findall(r"\b" + re.escape(m), f))
            if n:
                c[m] = n; tot[m] += n
        n_inst = len(re.findall(r"/\*[0-9a-f]{4}\*/", f))
        if c.get("UTCHMMA") or c.get("UTMALDG") or c.get("LDTM") or c.get("REDUX") or c.get("MATCH"):
            rows.append((name, n_inst, dict(c)))
    print("SASS mnemonic counts of see-vcn_b200/csrc/libseevcn_b200.so (cuobjdump -sass, sm_100a), produced by tools/sass_grep.py")
    print("UTCHMMA = tcgen05.mma (bf16), UTMALDG / UTMASTG = TMA tensor load / store, FENCE.VIEW.ASYNC = fence.proxy.async, LDTM/STTM = tcgen05.ld/st (tensor memory), UTCBAR = tcgen05.commit,")
    print("SYNCS = mbarrier ops, REDUX/MATCH = warp reduce / match, HMMA = legacy mma.sync (expected: 0).\n")
    print("whole library: " + ", ".join(f"{k} {v}" for k, v in sorted(tot.items())) + "\n")
    for name, n, c in rows:
        print(f"{name[:110]}\n    {n} instructions; " + ", ".join(f"{k} {v}" for k, v in sorted(c.items())))


if __name__ == "__main__":
    main()
