#!/bin/bash
# usage: tools/ab_lib.sh <old.so> [bench args...]   (on the GPU box)  A/B of two builds of the library on one box:
# runs bench.py with the in-tree library, then with <old.so> copied over it, and prints the per-stage tables.
set -u
old=$1; shift
L=see-vcn_b200/csrc/libseevcn_b200.so
show() { python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$1', round(d['value'],1), d['unit'], round(d['ms_per_step'],3), 'ms/step  e2e', round(d['e2e']['value'],1))
print('   ', [(s['group'], round(s['ms_per_step'],3)) for s in d['stages']])"; }
cp $L /tmp/new.so
for rep in 1 2; do
  cp /tmp/new.so $L; timeout 300 python bench.py --no-cpu "$@" 2>/dev/null | show new
  cp $old $L;        timeout 300 python bench.py --no-cpu "$@" 2>/dev/null | show old
done
cp /tmp/new.so $L
