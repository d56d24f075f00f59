#!/bin/bash
# bring-up build of the C-ABI library with in-kernel phase timers (never used by the product path / tests / bench)
set -e
cd "$(dirname "$0")/.."
mkdir -p tools/_dbg
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -shared -Xcompiler -fPIC --threads 4 -DAV2V_ATTN_TIMERS -DAV2V_GEMM_BRINGUP \
  -o tools/_dbg/libanyv2v_b200_timers.so anyv2v_b200/csrc/abi.cu anyv2v_b200/csrc/elementwise.cu anyv2v_b200/csrc/groupnorm.cu anyv2v_b200/csrc/gemm_tcgen05.cu anyv2v_b200/csrc/attention_tcgen05.cu anyv2v_b200/csrc/attention2q_tcgen05.cu anyv2v_b200/csrc/attention_tfused_tcgen05.cu
