#!/bin/bash
# old build vs new build: bitwise comparison, then the GPU test suite and the epilogue timing
mkdir -p gpurun_out
JNE_LIBRARY=$PWD/johansen_null_eigenspectra_b200/libjne_exp_old.so python tools/dump_eigs.py /tmp/old.npz 2>&1 | tail -2
python tools/dump_eigs.py /tmp/new.npz 2>&1 | tail -2
python tools/cmp_dumps.py /tmp/old.npz /tmp/new.npz 2>&1 | tail -15 | tee gpurun_out/regress_cmp.txt
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -6
python tools/exp_epi_multi.py 2>&1 | tee gpurun_out/exp_epi_multi_new.txt
