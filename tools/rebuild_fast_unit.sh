#!/bin/bash
# Recompile only the WCNS5-JS fast translation unit with extra flags and relink: tools/rebuild_fast_unit.sh -DHB2_SKEW=1
# (A/B builds of the headline kernels without the 4-minute full build; `python -m hamers_b200.build --force` restores the default)
set -e
cd "$(dirname "$0")/.."
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -I include \
    -DHB2_MATH=1 -fmad=true "$@" -c hamers_b200/csrc/hb2_sweeps.cu -o build/hb2_sweeps_fast.o
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o hamers_b200/libhamers_b200.so \
    build/hb2_sweeps_exact.o build/hb2_sweeps_exact_z.o build/hb2_sweeps_exact_ld.o build/hb2_sweeps_fast.o \
    build/hb2_sweeps_fast_z.o build/hb2_sweeps_fast_ld.o build/hb2_abi.o build/hb2_diffusive.o build/hb2_amr.o build/hb2_level.o
echo relinked with "$@"
