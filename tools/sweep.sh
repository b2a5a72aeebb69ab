#!/bin/sh
# developer tool: rebuild with one constant changed and print the per-stage times (run on the GPU box)
F=weiner_slamit_v2_b200/csrc/orb_extractor.cu
cp $F /tmp/orig.cu
run() { sh weiner_slamit_v2_b200/csrc/build.sh 2>&1 | grep -E "error" ; echo "$1: $(timeout 200 python tools/gpu_first.py 256 2>&1 | grep 'stage ms')"; cp /tmp/orig.cu $F; }
sed -i 's/constexpr int FAST_WARPS = 4;/constexpr int FAST_WARPS = 2;/' $F; run "FAST_WARPS=2"
sed -i 's/constexpr int FAST_WARPS = 4;/constexpr int FAST_WARPS = 8;/' $F; run "FAST_WARPS=8"
sed -i 's/constexpr int DESC_WARPS = 8;/constexpr int DESC_WARPS = 4;/' $F; run "DESC_WARPS=4"
sed -i 's/constexpr int DESC_WARPS = 8;/constexpr int DESC_WARPS = 16;/' $F; run "DESC_WARPS=16"
sed -i 's/constexpr int QT_THREADS = 256;/constexpr int QT_THREADS = 128;/' $F; run "QT_THREADS=128"
sed -i 's/constexpr int QT_THREADS = 256;/constexpr int QT_THREADS = 512;/' $F; run "QT_THREADS=512"
sed -i 's/(long long)h->numSMs \* 4);/(long long)h->numSMs * 3);/' $F; run "BLUR ctas/SM=3"
sed -i 's/(long long)h->numSMs \* 4);/(long long)h->numSMs * 8);/' $F; run "BLUR ctas/SM=8 (smem allows 4)"
sed -i 's/constexpr int BL_ROWS = 36, BL_WARPS = 4;/constexpr int BL_ROWS = 36, BL_WARPS = 2;/; s/(long long)h->numSMs \* 4);/(long long)h->numSMs * 8);/' $F; run "BL_WARPS=2 x8"
sh weiner_slamit_v2_b200/csrc/build.sh
