#!/bin/bash
# Development aid: build a variant of the library with extra nvcc flags (or from another git revision) next to the product .so
#   tools/build_variant.sh NAME "EXTRA_FLAGS" [GIT_REV]      ->  gaudi_b200/csrc/lib_NAME.so   (select it with GAUDI_B200_LIB)
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
NAME=$1; EXTRA=$2; REV=$3
W=/tmp/gb_variant_$NAME
rm -rf $W; mkdir -p $W/gaudi_b200 $W/include
if [ -n "$REV" ]; then
  git -C $ROOT archive $REV gaudi_b200/csrc include | tar -x -C $W
else
  cp -r $ROOT/gaudi_b200/csrc $W/gaudi_b200/; cp $ROOT/include/*.h $W/include/
fi
rm -f $W/gaudi_b200/csrc/*.o $W/gaudi_b200/csrc/*.so
make -C $W/gaudi_b200/csrc -j 8 EXTRA="$EXTRA" > $W/build.log 2>&1 || { tail -20 $W/build.log; exit 1; }
cp $W/gaudi_b200/csrc/libgaudi_b200.so $ROOT/gaudi_b200/csrc/lib_$NAME.so
echo "built gaudi_b200/csrc/lib_$NAME.so"
