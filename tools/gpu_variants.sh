#!/bin/bash
# headline (and five-eqn with FE=1) A/B over tuning variants built by tools/build_variant.py:  bash tools/gpu_variants.sh TAG v1 v2 ...
TAG=$1; shift
args=("A=product")
for v in "$@"; do args+=("HAMERS_B200_LIB=$PWD/hamers_b200/libhamers_b200_$v.so"); done
bash tools/gpu_ab.sh $TAG "${args[@]}"
if [ "${FE:-0}" = "1" ]; then for v in "$@"; do bash tools/gpu_fe_ab.sh ${TAG}_fe $v | tail -1; done; fi
