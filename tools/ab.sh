#!/bin/bash
# A/B of two builds of the library on the same box: tools/ab.sh "<command>" [rounds]   (ab/libA.so = baseline, ab/libB.so = candidate)
cmd="$1"; rounds="${2:-3}"
for r in $(seq 1 "$rounds"); do
  for v in A B; do
    cp ab/lib$v.so canny_edge_b200/libcanny_b200.so
    echo "== $v round $r"; eval "$cmd"
  done
done
