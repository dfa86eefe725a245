#!/bin/bash
# Experiment aid: build a variant of the library with extra -D switches into its OWN file (never the product library):
#   tools/build_variant.sh exp1 -DK4A_STAGGER_NS=1500   ->  dmdqn_b200/libdmdqn_b200_exp1.so
# and run against it with DMDQN_PROFILING_LIB=libdmdqn_b200_exp1.so.
set -e
name=$1; shift
cd "$(dirname "$0")/../dmdqn_b200/csrc"
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 --expt-extended-lambda -Xcompiler -fPIC -shared "$@" \
     -o ../libdmdqn_b200_$name.so api.cu featurize.cu act.cu replay.cu learn.cu learn_tc.cu
echo ../libdmdqn_b200_$name.so
