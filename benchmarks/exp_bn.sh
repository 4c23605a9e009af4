# tile width chosen by the cost model vs forced (VIP_GEMM_FORCE_BN) on the deep-K GCViT / ResNet-RS shapes
for s in "200704 1152 384 ln" "200704 768 384 ln gelu" "200704 384 384 res lo" "200704 384 768 res lo" "50176 2304 768 ln" "50176 1536 768 ln gelu" "173056 256 1024 relu" "43904 512 2048 relu"; do
  for bn in 0 128 256; do
    VIP_GEMM_FORCE_BN=$bn python benchmarks/one_gemm.py $s | sed "s/^/bn=$bn /"
  done
done
