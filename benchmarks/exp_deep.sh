# shallow (two CTAs per SM, BN <= 128) vs deep (one CTA per SM) configuration for the K = 384 / 768 contractions
for kb in 4 6 12; do
  for s in "200704 1152 384 ln" "200704 768 384 ln gelu" "200704 384 384" "802816 576 192 ln"; do
    VIP_GEMM_DEEP_KB=$kb python benchmarks/one_gemm.py $s | sed "s/^/deep_kb=$kb /"
  done
done
