#!/usr/bin/env python
"""Offline instruction budget of a kernel's hot loop: dump the SASS of one function of libcsgpu.so and
print the instruction count between consecutive occurrences of a marker opcode (e.g. MUFU.LG2: one per
pixel in K3's log path), plus an opcode histogram of that stretch.
usage: scripts/sass_loop.py <function substring> <marker opcode> [--show N]"""
import collections
import re
import subprocess
import sys

LIB = "configurable_spectrograms_b200/libcsgpu.so"


def main():
    fn, marker = sys.argv[1], sys.argv[2]
    show = int(sys.argv[sys.argv.index("--show") + 1]) if "--show" in sys.argv else -1
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout.splitlines()
    starts = [i for i, l in enumerate(out) if "Function :" in l]
    hits = [i for i in starts if fn in out[i]]
    if not hits:
        sys.exit(f"no function matching {fn!r}")
    for h in hits:
        end = next((s for s in starts if s > h), len(out))
        ins = [re.sub(r"/\*[0-9a-fx]+\*/", "", l).strip() for l in out[h:end] if re.search(r"/\*[0-9a-f]{4}\*/", l)]
        ins = [re.sub(r"\s*;\s*$", "", i) for i in ins]
        print(out[h].strip()[:140], f"-- {len(ins)} instructions")
        marks = [i for i, s in enumerate(ins) if marker in s.split()[0:2] or s.startswith(marker) or f" {marker}" in s[:24]]
        gaps = [b - a for a, b in zip(marks[:-1], marks[1:])]
        print(f"  {len(marks)} x {marker}; gaps: {gaps}")
        if show >= 0 and show + 1 < len(marks):
            seg = ins[marks[show] : marks[show + 1]]
            hist = collections.Counter(s.split()[1].split(".")[0] if s.startswith("@") else s.split()[0].split(".")[0] for s in seg)
            print("  ", dict(hist.most_common()))
            for s in seg:
                print("     ", s[:110])


if __name__ == "__main__":
    main()
