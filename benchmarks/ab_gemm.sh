# A/B of two builds of the library on the same box, interleaved: bash benchmarks/ab_gemm.sh <other .so>
other=$1
while read -r shape; do
  for rep in 1 2; do
    VIP_LIB_PATH=$other python benchmarks/one_gemm.py $shape | sed "s/^/base /"
    python benchmarks/one_gemm.py $shape | sed "s/^/new  /"
  done
done <<'S'
200704 384 384
200704 384 384 res lo
200704 1152 384 ln
200704 768 384 ln gelu
802816 192 192 res lo
802816 576 192 ln
3211264 96 96 res lo
3211264 288 96 ln
173056 1024 256 res
2560000 256 64 res
2560000 64 256 relu
S
