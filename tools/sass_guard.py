#!/usr/bin/env python
"""Which kernels changed since the library was last validated on a GPU?

    python tools/sass_guard.py record   # after a green `pytest -m gpu` run: store a hash of every kernel's SASS
    python tools/sass_guard.py check    # before shipping a rebuild without GPU time: list kernels whose SASS differs

The hash covers the instruction stream only (addresses, encodings and -lineinfo line comments are stripped), so moving
code around a file or adding a kernel next to a validated one does not flag it; any change in generated code does.
Reads libitcpd_b200.so with cuobjdump; the manifest lives in profiles/sass_manifest.json."""
import hashlib
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "itensorcpd.jl_b200", "lib", "libitcpd_b200.so")
MANIFEST = os.path.join(ROOT, "profiles", "sass_manifest.json")


def kernel_hashes():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    hashes, name, h = {}, None, None
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            if name:
                hashes[name] = h.hexdigest()
            name, h = m.group(1), hashlib.sha1()
            continue
        if name is None or "//##" in line:
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(.*?);", line)   # "/*0010*/  S2R R7, SR_TID.X ;  /* encoding */"
        if m:
            h.update(re.sub(r"\s+", " ", m.group(1)).encode())
    if name:
        hashes[name] = h.hexdigest()
    return hashes


def demangle(names):
    try:
        out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True, check=True).stdout.splitlines()
        return dict(zip(names, out))
    except Exception:
        return {n: n for n in names}


def main():
    mode = sys.argv[1] if len(sys.argv) > 1 else "check"
    cur = kernel_hashes()
    if mode == "record":
        json.dump({"note": "SASS hashes of the kernels in the library build that passed pytest -m gpu on a B200", "kernels": cur},
                  open(MANIFEST, "w"), indent=1, sort_keys=True)
        print(f"recorded {len(cur)} kernels -> {MANIFEST}")
        return 0
    old = json.load(open(MANIFEST))["kernels"]
    names = demangle(sorted(set(cur) | set(old)))
    changed = [n for n in cur if n in old and old[n] != cur[n]]
    added = [n for n in cur if n not in old]
    removed = [n for n in old if n not in cur]
    for tag, lst in (("CHANGED", changed), ("new", added), ("removed", removed)):
        for n in sorted(lst):
            print(f"{tag:8s} {names[n][:150]}")
    print(f"{len(cur) - len(changed) - len(added)} of {len(cur)} kernels identical to the validated build; {len(changed)} changed, {len(added)} new")
    return 1 if changed else 0


if __name__ == "__main__":
    sys.exit(main())
