"""Per-kernel digest of the SASS in libmap2d_b200.so (instruction text only, addresses and encodings stripped):
    python scripts/sass_digest.py > /tmp/before.txt ; <edit, rebuild> ; python scripts/sass_digest.py | diff /tmp/before.txt -
A kernel whose digest is unchanged executes the same instructions: refactors of shared device code can be checked on a
box without a GPU."""
import hashlib
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "pi-slam-fusion_b200", "libmap2d_b200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
cur, body = None, {}
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        body[cur] = []
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4}\*/\s+(.*?)\s*/\*", line)
    if cur and m:
        body[cur].append(m.group(1))
for k in sorted(body):
    print(k, len(body[k]), hashlib.sha1("\n".join(body[k]).encode()).hexdigest()[:16])
