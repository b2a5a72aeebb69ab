# developer tool (run on the GPU box): k_blur with other tile heights / blocks per SM, per-stage times of each build
F=weiner_slamit_v2_b200/csrc/orb_extractor.cu
cp $F /tmp/orig.cu
run() { sh weiner_slamit_v2_b200/csrc/build.sh 2>&1 | grep -E " error" ; echo "$1: $(timeout 200 python tools/stage_times.py 256 2>&1 | tail -1)"; cp /tmp/orig.cu $F; }
sed -i 's/constexpr int BL_ROWS = 24, BL_WARPS = 4;/constexpr int BL_ROWS = 18, BL_WARPS = 4;/; s/constexpr int BL_CTAS_PER_SM = 5; /constexpr int BL_CTAS_PER_SM = 7; /' $F; run "rows18 x7"
sed -i 's/constexpr int BL_ROWS = 24, BL_WARPS = 4;/constexpr int BL_ROWS = 18, BL_WARPS = 4;/; s/constexpr int BL_CTAS_PER_SM = 5; /constexpr int BL_CTAS_PER_SM = 6; /' $F; run "rows18 x6"
sed -i 's/constexpr int BL_ROWS = 24, BL_WARPS = 4;/constexpr int BL_ROWS = 30, BL_WARPS = 4;/; s/constexpr int BL_CTAS_PER_SM = 5; /constexpr int BL_CTAS_PER_SM = 4; /' $F; run "rows30 x4"
sed -i 's/constexpr int BL_ROWS = 24, BL_WARPS = 4;/constexpr int BL_ROWS = 48, BL_WARPS = 4;/; s/constexpr int BL_CTAS_PER_SM = 5; /constexpr int BL_CTAS_PER_SM = 3; /' $F; run "rows48 x3"
sh weiner_slamit_v2_b200/csrc/build.sh
