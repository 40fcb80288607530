#!/bin/bash
# Installs the UNMODIFIED reference package (bglid/GTCRN-Micro, /root/reference) into the git-ignored baseline/_ref so that
# `bench.py --impl reference` can run the reference's own GTCRNMicro / StreamGTCRNMicro on a GPU box, where /root/reference does
# not exist.  Dev container only.  Nothing under baseline/_ref is committed (see .gitignore); it travels with gpurun snapshots.
#
# pip cannot build the tree as it stands ("Multiple top-level packages discovered in a flat-layout: assets, esp32_s3,
# gtcrn_micro", and /root/reference is read-only), so the install runs from a copy under /tmp that holds pyproject.toml, README.md
# and the gtcrn_micro/ package only -- no source file is edited.  Dependencies are not resolved (--no-deps): the hot path needs
# torch, numpy and einops, which the image has.  The checkpoint (package data that setuptools does not pick up) is copied next
# to the installed package, where infer.py's default config expects it relative to the package root.
set -e
cd "$(dirname "$0")/.."
REF=${GTCRN_REFERENCE:-/root/reference}
[ -d "$REF/gtcrn_micro" ] || { echo "reference not found at $REF"; exit 1; }
TMP=$(mktemp -d /tmp/gtcrn_ref_XXXX)
cp "$REF/pyproject.toml" "$REF/README.md" "$TMP/"
cp -r "$REF/gtcrn_micro" "$TMP/gtcrn_micro"
rm -rf baseline/_ref
python -m pip install -q --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse --target baseline/_ref "$TMP"
mkdir -p baseline/_ref/gtcrn_micro/ckpts
cp "$REF/gtcrn_micro/ckpts/best_model_dns3.tar" baseline/_ref/gtcrn_micro/ckpts/
rm -rf "$TMP"
( cd baseline/_ref && for f in gtcrn_micro/models/gtcrn_micro.py gtcrn_micro/streaming/gtcrn_micro_stream.py gtcrn_micro/streaming/conversion/convert.py \
    gtcrn_micro/streaming/conversion/convolution.py; do cmp "$f" "$REF/$f" || { echo "MISMATCH $f"; exit 1; }; done )
echo "installed: $(ls baseline/_ref)"
