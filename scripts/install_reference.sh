#!/bin/bash
# Installs the UNMODIFIED reference into baseline/_ref (git-ignored, travels to the GPU box) for `bench.py --impl reference`.
# The reference ships no setup.py / pyproject.toml, so pip has nothing to build from the read-only checkout: the checkout
# is copied to a scratch directory, a packaging-only setup.py is written THERE (no reference source is touched or copied
# into this repo's history), and pip installs from that copy.
set -e
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
REF="${1:-/root/reference}"
TMP="$(mktemp -d /tmp/dfd_ref_XXXX)"
cp -r "$REF/." "$TMP/"
cat > "$TMP/setup.py" <<'PY'
from setuptools import setup
import glob, os
pkgs = [d for d in os.listdir('.') if os.path.isdir(d) and not d.startswith('.') and d not in ('build', 'dist')
        and not d.endswith('.egg-info')]
setup(name='dfd-starter-reference', version='0.0.0',
      packages=[r.replace(os.sep, '.') for p in pkgs for r, _, fs in os.walk(p) if any(f.endswith('.py') for f in fs)],
      py_modules=[f[:-3] for f in glob.glob('*.py') if f != 'setup.py'],
      include_package_data=True, package_data={'': ['*.txt', '*.proto', '*.json']})
PY
rm -rf "$ROOT/baseline/_ref"
python -m pip install --no-index --no-build-isolation --find-links /opt/wheelhouse --no-deps \
    --target "$ROOT/baseline/_ref" "$TMP"
rm -rf "$TMP"
ls "$ROOT/baseline/_ref"
