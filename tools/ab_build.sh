#!/bin/bash
# Build the library of another commit as gansformer-reproducibility-challenge_b200/libgf_attn_<tag>.so for same-box A/B runs:
#   tools/ab_build.sh <commit> <tag>;  then  GF_ATTN_LIB=$PWD/gansformer-reproducibility-challenge_b200/libgf_attn_<tag>.so python tools/attn_bench.py
set -e
commit=$1; tag=$2
root=$(git rev-parse --show-toplevel)
tmp=$(mktemp -d)
git -C "$root" archive "$commit" gansformer-reproducibility-challenge_b200/csrc include | tar -x -C "$tmp"
srcs=$(ls "$tmp"/gansformer-reproducibility-challenge_b200/csrc/*.cu)
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -shared -o "$root/gansformer-reproducibility-challenge_b200/libgf_attn_$tag.so" $srcs
rm -rf "$tmp"
echo "built libgf_attn_$tag.so from $commit"
