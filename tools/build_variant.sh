#!/bin/bash
# Dev tool: build a variant of the library under build/variants/<name>.so.
#   tools/build_variant.sh <name> [git-rev|-] [extra nvcc flags...]    ("-" = working tree)
set -e
cd "$(dirname "$0")/.."
name=$1; rev=${2:--}; shift; shift || true
src=squishrs_b200
if [ "$rev" != "-" ]; then rm -rf /tmp/sqv_$name; mkdir -p /tmp/sqv_$name; git archive "$rev" squishrs_b200 include | tar -x -C /tmp/sqv_$name; src=/tmp/sqv_$name/squishrs_b200; inc=/tmp/sqv_$name/include; else inc=include; fi
mkdir -p build/variants /tmp/sqv_obj_$name
objs=""
for f in $src/csrc/*.cu $src/host/*.cpp; do
  o=/tmp/sqv_obj_$name/$(basename $f).o
  x=""; case $f in *.cpp) x="-x cu";; esac
  nvcc $x -gencode arch=compute_100a,code=sm_100a "$@" -O3 -lineinfo -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr -I $inc -I $src/csrc -c $f -o $o &
  objs="$objs $o"
done
wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o build/variants/$name.so $objs -lcudart -lpthread
echo built build/variants/$name.so
