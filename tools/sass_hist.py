"""Static SASS opcode histogram of selected kernels (cuobjdump -sass; no GPU needed).
Usage: python tools/sass_hist.py <object or .so> <kernel name substring> [...]"""
import collections
import re
import subprocess
import sys


def main():
    obj, pats = sys.argv[1], sys.argv[2:]
    txt = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    hist, cur = collections.defaultdict(collections.Counter), None
    for line in txt.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
        if m and cur:
            op = m.group(1)
            key = op.split(".")[0] + (".WIDE" if ".WIDE" in op else "") + (".HI" if ".HI" in op else "")
            hist[cur][key] += 1
    for k, h in hist.items():
        if any(p in k for p in pats):
            tot = sum(h.values())
            dem = subprocess.run(["c++filt", k], capture_output=True, text=True).stdout.strip()[:110]
            print(f"{dem}\n  static SASS instructions: {tot}")
            for op, c in h.most_common(12):
                print(f"    {op:12s} {c:6d}  {100 * c / tot:5.1f} %")


if __name__ == "__main__":
    main()
