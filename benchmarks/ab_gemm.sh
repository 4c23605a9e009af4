# A/B of two builds of the library on the same box, interleaved: bash benchmarks/ab_gemm.sh <other .so>
other=$1
while read -r shape; do
  for rep in 1 2; do
    VIP_LIB_PATH=$other python benchmarks/one_gemm.py $shape | sed "s/^/base /"
    python benchmarks/one_gemm.py $shape | sed "s/^/new  /"
  done
done <<'S'
173056 1024 256 se
692224 512 128 se
2560000 256 64 se
2560000 64 256 relu
802816 576 192 ln
802816 192 192 res lo
3211264 288 96 ln
3211264 96 96 res lo
50176 768 256 ln
50176 256 256 res lo
S
