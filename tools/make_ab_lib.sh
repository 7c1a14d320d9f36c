#!/bin/bash
# Build the library of another commit into tools/_ab/libemdr2_old.so for the interleaved A/B timing tools
# (gpu_ab_gemm.py, gpu_ab_attn.py, gpu_ab_sustained.py).  Usage: tools/make_ab_lib.sh <commit> [name.so]
set -e
REV=${1:?commit}
OUT=${2:-libemdr2_old.so}
ROOT=$(cd "$(dirname "$0")/.." && pwd)
TMP=$(mktemp -d)
git -C "$ROOT" worktree add -f "$TMP/wt" "$REV" > /dev/null
(cd "$TMP/wt" && python emdr2_b200/build.py > /dev/null)
mkdir -p "$ROOT/tools/_ab"
cp "$TMP/wt/emdr2_b200/libemdr2_b200.so" "$ROOT/tools/_ab/$OUT"
git -C "$ROOT" worktree remove --force "$TMP/wt"
git -C "$ROOT" worktree prune
rm -rf "$TMP"
echo "$ROOT/tools/_ab/$OUT"
