#!/bin/bash
# Install the UNMODIFIED reference (facebookresearch/jepa) into baseline/_ref (git-ignored, travels with gpurun).
# The reference's setup.py declares no packages and the tree has no __init__.py files, so setuptools' auto-discovery
# treats src/ as a "src layout" and flattens it (src.models -> models), which breaks the reference's own
# `from src.models...` imports.  We therefore install from a copy under /tmp whose setup.py (packaging metadata ONLY -
# no module source is touched) lists the namespace packages explicitly.  Outcome is recorded in DESIGN.md section 9.
set -euo pipefail
REF=${1:-/root/reference}
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
TMP=$(mktemp -d /tmp/jepa_ref_XXXX)
cp -r "$REF"/. "$TMP"/
python - "$TMP/setup.py" <<'PY'
import sys
p = sys.argv[1]
s = open(p).read()
s = s.replace("from setuptools import setup", "from setuptools import setup, find_namespace_packages")
s = s.replace('python_requires=">=3.9",',
              'python_requires=">=3.9",\n        packages=find_namespace_packages(include=["src*", "app*", "evals*"]),')
open(p, "w").write(s)
PY
rm -rf "$ROOT/baseline/_ref"
mkdir -p "$ROOT/baseline"
(cd "$TMP" && python -m pip install --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse \
    --target "$ROOT/baseline/_ref" "$TMP" 2>&1 | tail -3)
rm -rf "$TMP"
# byte-identity check of every installed module against the reference tree
python - "$REF" "$ROOT/baseline/_ref" <<'PY'
import filecmp, os, sys
ref, dst = sys.argv[1], sys.argv[2]
n = bad = 0
for top in ("src", "app", "evals"):
    for d, _, fs in os.walk(os.path.join(dst, top)):
        for f in fs:
            if f.endswith(".py"):
                n += 1
                rel = os.path.relpath(os.path.join(d, f), dst)
                if not filecmp.cmp(os.path.join(d, f), os.path.join(ref, rel), shallow=False):
                    bad += 1
                    print("DIFFERS", rel)
print(f"baseline/_ref: {n} modules installed, {bad} differ from {ref}")
sys.exit(1 if bad or n == 0 else 0)
PY
