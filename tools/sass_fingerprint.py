#!/usr/bin/env python3
"""One md5 per kernel of libtensorkrylov_b200.so over its SASS instruction text (addresses and encodings stripped).
Lets a host-side change or a refactor prove that the device code that was measured is the device code that ships:

    python tools/sass_fingerprint.py                      # print
    python tools/sass_fingerprint.py --write              # refresh profiles/sass_fingerprint.txt
    python tools/sass_fingerprint.py --check              # exit 1 and list the kernels that differ from the file
"""
import argparse
import hashlib
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "tensorkrylov.jl_b200", "libtensorkrylov_b200.so")
REC = os.path.join(ROOT, "profiles", "sass_fingerprint.txt")


def fingerprint(lib=LIB):
    sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
    out, cur, buf = {}, None, []
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            if cur:
                out[cur] = hashlib.md5("\n".join(buf).encode()).hexdigest()
            cur, buf = m.group(1), []
        elif cur:
            t = re.sub(r"/\*[0-9a-fx]+\*/", "", line).strip()
            if t:
                buf.append(t)
    if cur:
        out[cur] = hashlib.md5("\n".join(buf).encode()).hexdigest()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--write", action="store_true")
    ap.add_argument("--check", action="store_true")
    a = ap.parse_args()
    fp = fingerprint()
    lines = [f"{h}  {name}" for name, h in sorted(fp.items())]
    if a.write:
        with open(REC, "w") as f:
            f.write("\n".join(lines) + "\n")
        print(f"{len(fp)} kernels -> {REC}")
        return 0
    if a.check:
        rec = dict(reversed(l.split("  ", 1)) for l in open(REC).read().splitlines() if l.strip())
        diff = sorted(set(k for k in fp if fp[k] != rec.get(k)) | set(rec) - set(fp))
        for k in diff:
            print("differs:", k)
        print(f"{len(fp)} kernels, {len(diff)} differ from {os.path.relpath(REC, ROOT)}")
        return 1 if diff else 0
    print("\n".join(lines))
    return 0


if __name__ == "__main__":
    sys.exit(main())
