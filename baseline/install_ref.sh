#!/bin/bash
# Installs the unmodified reference into baseline/_ref (git-ignored; it travels to the GPU box with the gpurun snapshot):
#   pip install --no-index --no-build-isolation --no-deps --target baseline/_ref <copy of /root/reference>
# --no-deps: monty / hydra-core / pytorch-lightning / torch==1.4.0 are not in the offline wheelhouse; the hot path needs none
# of them beyond monty.collections.AttrDict (baseline/ref_loader.py).  The copy under /tmp is there because setuptools writes
# build/ and *.egg-info into the source tree and /root/reference is read-only.
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
SRC="${1:-/root/reference}"
[ -d "$SRC/torch_scae" ] || { echo "install_ref: $SRC has no torch_scae package" >&2; exit 1; }
TMP="$(mktemp -d)"
cp -r "$SRC" "$TMP/ref"
rm -rf "$HERE/_ref"
python -m pip install --quiet --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse --target "$HERE/_ref" "$TMP/ref"
rm -rf "$TMP"
echo "installed $(ls "$HERE/_ref" | tr '\n' ' ')"
