V=gflow_b200/_lib/variants
run() { # name lib split
  GFB_BWD_SPLIT=$3 GFLOW_B200_LIB=$2 GFLOW_B200_NO_EXT=1 python tools/kernel_times.py fused 30 cfg2 synthetic flush 2>/dev/null | grep -E "blend_bwd|tile_sort_blend|us/step" | awk -v n="$1" '{print n": "$0}' | cut -c1-150
}
run old $V/libgfb_old.so 0
run base0 $V/libgfb_base.so 0
run base1 $V/libgfb_base.so 1
run s25 $V/libgfb_split25.so 1
run s35 $V/libgfb_split35.so 1
run s35b $V/libgfb_split35b.so 1
