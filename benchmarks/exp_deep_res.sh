# residual contractions (ResNet-RS conv_3 shapes at batch 1024): shallow two-CTA configuration (single staging-buffer set)
# vs the deep one (two sets: the next tile's residual is fetched a tile ahead); VIP_GEMM_DEEP_RES_KB = smallest K/64 kept shallow
for kb in 4 3 1 0; do
  for s in "173056 1024 256 res" "692224 512 128 res" "2560000 256 64 res" "50176 2048 512 res"; do
    VIP_GEMM_DEEP_RES_KB=$kb python benchmarks/one_gemm.py $s | sed "s/^/deep_res_kb=$kb /"
  done
done
