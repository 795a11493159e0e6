#!/bin/bash
# bitwise regression of the default build against the reference build, then timing of the listed variants
JNE_LIBRARY=$PWD/johansen_null_eigenspectra_b200/libjne_exp_old.so python tools/dump_eigs.py /tmp/old.npz 2>&1 | tail -1
python tools/dump_eigs.py /tmp/new.npz 2>&1 | tail -1
python tools/cmp_dumps.py /tmp/old.npz /tmp/new.npz 2>&1 | grep -v "_m3\|multi3" | tail -5
./tools/exp_lib_compare.sh "$@"
