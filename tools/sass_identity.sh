#!/bin/bash
# Is the SASS of the SHIPPED tokeniser instantiations unchanged between a git ref and the working tree?  (Build container, no GPU.)
# The opt-in variants are added as extra template parameters / `if constexpr` branches of rq_tc_kernel; this is the check that
# they did not perturb the code that runs by default and that the round's measurements were taken with.
#   bash tools/sass_identity.sh [ref=0ac61f5]      compares rq_tc_kernel<0|1 trace, vec, pair 0|1> with all later options off
set -e
cd "$(dirname "$0")/.."
REF=${1:-0ac61f5}
W=scratch/sass_identity; rm -rf $W; mkdir -p $W/ref
git archive $REF rq_vae_recommender_b200/csrc | tar -x -C $W/ref
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -cubin"
nvcc $FLAGS -o $W/ref.cubin $W/ref/rq_vae_recommender_b200/csrc/rq_tc.cu
nvcc $FLAGS -o $W/new.cubin rq_vae_recommender_b200/csrc/rq_tc.cu
dump() { cuobjdump -sass -fun "$2" "$1" | grep -E "^\s+/\*[0-9a-f]{4}\*/" | sed 's/\/\*[0-9a-f]*\*\/\s*$//'; }
names() { cuobjdump -sass "$1" | grep -o "_Z12rq_tc_kernel[A-Za-z0-9_]*" | sort -u; }
rc=0
for old in $(names $W/ref.cubin); do
  core=$(echo $old | sed -E 's/^_Z12rq_tc_kernelI(Lb[01]ELb[01]ELb[01]E).*/\1/')
  new=$(names $W/new.cubin | grep -E "^_Z12rq_tc_kernelI${core}(L[bi]0E)*Ev8TcParams$" | head -1)
  if [ -z "$new" ]; then echo "$old: no counterpart in the working tree"; rc=1; continue; fi
  n=$(diff <(dump $W/ref.cubin $old) <(dump $W/new.cubin $new) | wc -l)
  echo "$old -> $new: $(dump $W/new.cubin $new | wc -l) instructions, $n differing lines"
  [ "$n" = 0 ] || rc=1
done
exit $rc
