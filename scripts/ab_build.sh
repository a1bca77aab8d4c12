#!/bin/bash
# Developer aid: build the COMMITTED (HEAD) sources next to the working-tree build as
# py4cast_b200/lib/libpy4cast_b200_base.so, for A/B timing in one GPU session (P4C_LIB_AB=libpy4cast_b200_base.so).
set -e
cd "$(dirname "$0")/.."
rm -rf build/ab && mkdir -p build/ab/src
git archive HEAD py4cast_b200/csrc include | tar -x -C build/ab/src
make -s -C build/ab/src/py4cast_b200/csrc >/dev/null
cp build/ab/src/py4cast_b200/lib/libpy4cast_b200.so py4cast_b200/lib/libpy4cast_b200_base.so
ls -la py4cast_b200/lib/
