#!/bin/bash
# A/B build of the library with tuning defines: tools/mkvariant.sh <name> "<-D...>" [units to rebuild, default: constraints]
# Starts from the objects of the default build, recompiles the named units with the defines, links build/<name>/libcsg.so.
set -e
name=$1; extra=$2; shift 2; units=${@:-constraints}
cd "$(dirname "$0")/../certificate_stark_b200/csrc"
out=../../build/$name; mkdir -p $out/obj
cp -p ../lib/obj/*.o $out/obj/
for u in $units; do rm -f $out/obj/$u.o; done
make -s OUT=$out EXTRA="$extra" >/dev/null
ls -la $out/libcsg.so
