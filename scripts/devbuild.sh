#!/bin/bash
# quick development build of one shape: scripts/devbuild.sh [out.so]   (default /tmp/dev.so, shape (5,1,slack) only)
out=${1:-/tmp/dev.so}
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 --extended-lambda -Xcompiler -fPIC -shared \
  -DMPCB_DEV_SHAPE python-mpc_b200/csrc/mpc_b200.cu -o $out && \
cuobjdump --dump-resource-usage $out 2>/dev/null | grep -A1 -E "admm_tma_kernelId|ScaleOpd|FactorOpd|admm_check" | grep -E "Function|REG" | paste - - | sed -E 's/Function (_Z[A-Za-z0-9_]{0,60}).*REG:([0-9]+) STACK:([0-9]+).*/\1 REG \2 STACK \3/'
