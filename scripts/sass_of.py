#!/usr/bin/env python3
"""Print the SASS of one kernel of libsqg.so: scripts/sass_of.py <substring of mangled name> [lib]"""
import subprocess, sys
pat = sys.argv[1]
lib = sys.argv[2] if len(sys.argv) > 2 else "squigulator_b200/libsqg.so"
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
cur, keep = None, []
for line in out.splitlines():
    if "Function :" in line:
        cur = line.split("Function :")[1].strip()
    if cur and pat in cur:
        keep.append(line)
print("\n".join(keep))
