# developer tool (run on the GPU box): k_fast2 with its shared memory padded (fewer one-warp CTAs per SM), per-stage times of each build
F=weiner_slamit_v2_b200/csrc/orb_extractor.cu
cp $F /tmp/orig.cu
run() { sh weiner_slamit_v2_b200/csrc/build.sh 2>&1 | grep -E " error" ; echo "$1: $(timeout 200 python tools/stage_times.py 256 2>&1 | tail -1)"; cp /tmp/orig.cu $F; }
sed -i 's/h->fastSmem = P.fastLarge ? FastGeo2<38, 64>::SMEM_BYTES : FastGeo2<26, 42>::SMEM_BYTES;/h->fastSmem = (P.fastLarge ? FastGeo2<38, 64>::SMEM_BYTES : FastGeo2<26, 42>::SMEM_BYTES) + 2300;/' $F; run "fast smem +2300 (20 CTAs/SM)"
sed -i 's/h->fastSmem = P.fastLarge ? FastGeo2<38, 64>::SMEM_BYTES : FastGeo2<26, 42>::SMEM_BYTES;/h->fastSmem = (P.fastLarge ? FastGeo2<38, 64>::SMEM_BYTES : FastGeo2<26, 42>::SMEM_BYTES) + 1000;/' $F; run "fast smem +1000 (22 CTAs/SM)"
sh weiner_slamit_v2_b200/csrc/build.sh 2>&1 | grep " error"
