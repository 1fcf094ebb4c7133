#!/usr/bin/env python
"""Which kernels of two builds of one translation unit differ?  python tools/sass_diff.py old.o new.o
Compares the SASS instruction streams (addresses, encodings and line-info stripped) kernel by kernel, by demangled name
(the anonymous-namespace hash in the mangled names depends on the source path).  Used to show that a change to a shared
header leaves the kernels it should not touch instruction-for-instruction identical."""
import hashlib
import re
import subprocess
import sys


def signatures(obj):
    txt = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True, check=True).stdout
    out = {}
    for part in re.split(r"\n\s*Function : ", txt)[1:]:
        name, body = part.split("\n", 1)
        ins = [m.group(1).strip() for m in (re.match(r"\s*/\*[0-9a-f]{4}\*/\s+(.*?);", l) for l in body.split("\n")) if m]
        dem = subprocess.run(["c++filt", name.strip()], capture_output=True, text=True).stdout.strip()
        out[re.sub(r"hb2::\(anonymous namespace\)::", "", dem)] = (len(ins), hashlib.md5("\n".join(ins).encode()).hexdigest())
    return out


a, b = signatures(sys.argv[1]), signatures(sys.argv[2])
same = sorted(k for k in a if a[k] == b.get(k))
diff = sorted(set(a) | set(b) - set(same))
print(f"identical instruction streams: {len(same)}")
for k in same:
    print("   ", k)
print(f"different: {len([k for k in diff if k not in same])}")
for k in diff:
    if k not in same:
        print("   ", k, a.get(k, ("-",))[0], "->", b.get(k, ("-",))[0], "instructions")
