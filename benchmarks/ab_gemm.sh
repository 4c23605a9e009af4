# A/B of two builds of the library on the same box, interleaved: bash benchmarks/ab_gemm.sh <other .so>
other=$1
while read -r shape; do
  for rep in 1 2; do
    VIP_LIB_PATH=$other python benchmarks/one_gemm.py $shape | sed "s/^/base /"
    python benchmarks/one_gemm.py $shape | sed "s/^/new  /"
  done
done <<'S'
200704 384 384 res lo
200704 384 768 res lo
50176 768 768 res lo
50176 768 1536 res lo
200704 384 384
200704 384 768 relu
173056 256 2304 relu
43264 512 4608 relu
50176 2048 512 se
S
