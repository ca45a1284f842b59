#!/bin/bash
# A/B of kernel builds: production library against variants built beforehand into profiles/ab/ (same sources, another
# -D switch), same box, one process per measurement, twice, interleaved.
#   bash profiles/ab_div2.sh nodiv2 div4 ...   ->   profiles/ab/libotmb_<name>.so
P=oceantransportmatrixbuilder.jl_b200/libotmb.so
cp $P /tmp/prod.so
for rep in 1 2; do
  cp /tmp/prod.so $P; echo "== production"; timeout 200 python profiles/kbench.py BASE=1 2>&1 | tail -2
  for v in "$@"; do
    cp profiles/ab/libotmb_$v.so $P; echo "== $v"; timeout 200 python profiles/kbench.py BASE=1 2>&1 | tail -2
  done
done
cp /tmp/prod.so $P
