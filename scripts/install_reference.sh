#!/usr/bin/env bash
# Install the UNMODIFIED reference (node2vec-fugue 0.3.5, pure Python) into baseline/_ref so that
# `bench.py --impl reference` and the cpu_baseline leg can import its own functions on the GPU box
# (/root/reference does not exist there; baseline/_ref is git-ignored but travels with gpurun).
# Offline, no dependency resolution: the walk path of the reference needs only pandas
# (node2vec/randomwalk.py:1-14); fugue / gensim / pyspark are not installable here.
set -euo pipefail
REF="${1:-/root/reference}"
HERE="$(cd "$(dirname "$0")/.." && pwd)"
TMP="$(mktemp -d)"
cp -r "$REF" "$TMP/ref"                      # the source tree is read-only; pip builds in a copy
rm -rf "$HERE/baseline/_ref"
python -m pip install --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse \
    --target "$HERE/baseline/_ref" "$TMP/ref"
rm -rf "$TMP"
ls "$HERE/baseline/_ref/node2vec"
