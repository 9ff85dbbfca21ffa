#!/bin/bash
# Stage breakdown of an ncu capture of rt_wavefront_kernel: phases are found from the marker comments in the source.
# usage: tools/phases.sh <rep> [kernel-substr]
REP=$1; K=${2:-rt_wavefront_kernelILi3ELj0ELb0}
F=rust-raytracer_b200/csrc/rtb200_wavefront.cu
ln() { grep -n "$1" $F | head -1 | cut -d: -f1; }
SETUP=$(ln "stage the scene into shared memory"); REGEN=$(ln "auto regenerate = "); FILL=$(ln "regenerate(true, (uint32_t)tid)")
CH=$(ln "=== closest-hit ==="); CONST=$(ln "per-ray constants in the recentred"); TRAV=$(ln "warp-cooperative traversal ----")
NODE=$(ln "node step: lane"); LEAF=$(ln "leaf step: lane"); EXACT=$(ln "exact step: lane"); BRUTE=$(ln "MODE_BRUTE: hit_world")
MERGE=$(ln "if (alive) {"); SORT=$(ln "=== sort: compact"); BA=$(ln "// A: class counts"); BB=$(ln "// B: perm complete")
SHADE=$(ln "=== shade + regenerate"); RG=$(ln "regenerate(active && done"); BC=$(ln "__syncthreads_or(still_alive"); STATS=$(ln "statistics: one atomic per warp")
python tools/ncu_phases.py $REP rust-raytracer_b200/librtb200.so $K $F \
  setup:$SETUP-$((REGEN-1)) regen:$REGEN-$((FILL+2)) raysetup:$CH-$((TRAV-1)) trav_ctl:$TRAV-$((NODE-1)) node:$NODE-$((LEAF-1)) leaf:$LEAF-$((EXACT-1)) \
  exact:$EXACT-$((BRUTE-1)) merge:$MERGE-$((SORT-1)) sort:$SORT-$((BA-1)) syncA:$BA-$BA sort2:$((BA+1))-$((BB-1)) syncB:$BB-$BB shade:$SHADE-$((RG-1)) \
  regen_call:$RG-$((BC-1)) syncC:$BC-$((BC+6)) tail:$STATS-$((STATS+30))
