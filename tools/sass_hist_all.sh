#!/bin/bash
# Static SASS opcode histograms of the hot kernels + the Blackwell mnemonics of the whole library -> profiles/<round>/sass_hist_kernels.txt
# (runs without a GPU: cuobjdump on the in-tree libsdpcutsel.so).   bash tools/sass_hist_all.sh r02
RND=${1:-r02}
LIB=sdpcutsel-via-nn_b200/libsdpcutsel.so
OUT=profiles/$RND/sass_hist_kernels.txt
{
echo "# static SASS opcode histograms (cuobjdump -sass $LIB; tools/sass_hist.py), final build of round ${RND#r0}"
for k in k_mlp_i8ILi4ELi0ELi7ELb0 k_mlp_i8ILi4ELi0ELi4ELb0 k_prep_i8ILi5ELi7ELb1 k_prep_i8ILi5ELi7ELb0 k_score_feasILi5 k_score_nnILi5 k_sel_histILi3ELb1 k_rank_sort k_merge_rows; do
  echo; echo "== $k"; python tools/sass_hist.py $LIB $k
done
echo; echo "== Blackwell-native evidence over the whole library (mnemonic: count)"
cuobjdump -sass $LIB | grep -oE "\b(UTCIMMA|UTCHMMA|UTCQMMA|LDTM|STTM|UBLKCP|UTMALDG|UTCBAR|USETMAXREG|UTCATOMSWS|DMMA\.8x8x4|SYNCS\.PHASECHK|NANOSLEEP\.SYNCS)\b" | sort | uniq -c
} > $OUT
tail -12 $OUT
