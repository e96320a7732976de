#!/usr/bin/env bash
# Builds the instrumented library and prints cycles per timestep per phase (run under gpurun).
set -e
cd "$(dirname "$0")/../.."
nvcc -gencode arch=compute_100a,code=sm_100a -O3 --fmad=false -lineinfo -std=c++17 -Xcompiler -fPIC -shared \
  -DB200LP_PHASE_TIMING -o pywr_b200/libb200lp_timing.so pywr_b200/csrc/b200lp.cu
python benchmarks/tools/phase_timing.py
