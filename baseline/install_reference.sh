#!/usr/bin/env bash
# Installs the UNMODIFIED reference into baseline/_ref (git-ignored; travels to the GPU box with the gpurun snapshot).
#
# 1. the sanctioned offline install — fails in this image: the PEP-517 backend `hatchling` is not installed / not in
#    /opt/wheelhouse, and pyproject pins `requires-python <3.11` (the image has 3.12);
# 2. fallback: the package is pure Python, so a byte-for-byte copy of /root/reference/fl4health is the same artefact a
#    wheel would unpack.  Its third-party imports (flwr, torchmetrics, opacus) are satisfied by baseline/stubs/.
set -u
cd "$(dirname "$0")/.."
SRC="${FL4H_REFERENCE_SRC:-/root/reference}"
if python -m pip install --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse \
        --target baseline/_ref "$SRC" 2> baseline/_ref_pip_install.log; then
    echo "pip install succeeded"
else
    echo "pip install failed (see baseline/_ref_pip_install.log); copying the pure-Python package instead"
    mkdir -p baseline/_ref
    rm -rf baseline/_ref/fl4health
    cp -r "$SRC/fl4health" baseline/_ref/fl4health
    find baseline/_ref -name __pycache__ -type d -prune -exec rm -rf {} +
fi
diff -r -q -x __pycache__ "$SRC/fl4health" baseline/_ref/fl4health && echo "baseline/_ref/fl4health is identical to $SRC/fl4health"
