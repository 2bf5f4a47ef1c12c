#!/bin/bash
# tools/mkvariant.sh <name> [extra nvcc flags...]  -> variants/<name>.so built from the working tree (A/B runs, see tools/ab_step.py)
set -e
cd "$(dirname "$0")/.."
name=$1; shift
SRC=${SRC_ROOT:-.}
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -shared -Xcompiler -fPIC -cudart shared --threads 2 "$@" \
  -o variants/$name.so $SRC/tvc_ai_b200/csrc/tvc_abi.cu $SRC/tvc_ai_b200/csrc/tvc_rollout.cu $SRC/tvc_ai_b200/csrc/tvc_replay.cu $SRC/tvc_ai_b200/csrc/tvc_curiosity.cu
echo built variants/$name.so
