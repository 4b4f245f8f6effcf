"""Per-function SASS comparison of two objects / shared libraries, ignoring register numbers (ptxas renames registers
from one compilation to the next even for identical source).  Used to show that a rebuild which adds opt-in kernels
leaves the kernels that were validated on the GPU untouched:

    python tools/sass_diff.py OLD.o NEW.o
"""
import re
import subprocess
import sys


def functions(path):
    txt = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True, check=True).stdout
    out, cur = {}, None
    for line in txt.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            out[cur] = []
            continue
        if cur is None or re.match(r"^\s*/\*[0-9a-f]{4}\*/\s+/\* 0x[0-9a-f]{16} \*/\s*$", line):
            continue
        l = re.sub(r"/\* 0x[0-9a-f]{16} \*/", "", line)
        l = re.sub(r"/\*[0-9a-f]{4,}\*/", "", l).strip()
        l = re.sub(r"\bU?R\d+\b", "R", l)
        l = re.sub(r"\bU?P\d\b", "P", l)
        if l:
            out[cur].append(l)
    return out


def main():
    old, new = functions(sys.argv[1]), functions(sys.argv[2])
    same = 0
    for name, body in new.items():
        if name not in old:
            print(f"NEW      {name}")
        elif old[name] != body:
            print(f"CHANGED  {name}  ({len(old[name])} -> {len(body)} instructions)")
        else:
            same += 1
    for name in old:
        if name not in new:
            print(f"REMOVED  {name}")
    print(f"{same} functions identical modulo register names")


if __name__ == "__main__":
    main()
