#!/bin/bash
# A/B of builds of the library on the same box: [AB_LIBS="A B ..."] tools/ab.sh "<command>" [rounds]   (ab/lib<NAME>.so; default A = baseline, B = candidate)
cmd="$1"; rounds="${2:-3}"
for r in $(seq 1 "$rounds"); do
  for v in ${AB_LIBS:-A B}; do
    cp ab/lib$v.so canny_edge_b200/libcanny_b200.so
    echo "== $v round $r"; eval "$cmd"
  done
done
