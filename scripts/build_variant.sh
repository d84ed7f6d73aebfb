#!/bin/bash
# build a variant of libcgfd3d_b200.so with extra nvcc flags into cgfd3d_b200/variants/lib_<name>.so
# usage: scripts/build_variant.sh <name> "<extra nvcc flags>"
set -e
NAME=$1; FLAGS=$2
ROOT=$(cd $(dirname $0)/.. && pwd)
B=$ROOT/build/$NAME
mkdir -p $B/cgfd3d_b200/csrc $B/include $ROOT/cgfd3d_b200/variants
cp $ROOT/cgfd3d_b200/csrc/*.cu $ROOT/cgfd3d_b200/csrc/*.cuh $ROOT/cgfd3d_b200/csrc/*.h $ROOT/cgfd3d_b200/csrc/Makefile $B/cgfd3d_b200/csrc/
cp $ROOT/include/*.h $B/include/
make -s -j8 -C $B/cgfd3d_b200/csrc EXTRA="$FLAGS" OUT=$ROOT/cgfd3d_b200/variants/lib_$NAME.so
ls -la $ROOT/cgfd3d_b200/variants/lib_$NAME.so
