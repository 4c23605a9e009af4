# A/B of two builds of the library on the same box, interleaved: bash benchmarks/ab_gemm.sh <other .so>
other=$1
while read -r shape; do
  for rep in 1 2; do
    VIP_LIB_PATH=$other python benchmarks/one_gemm.py $shape | sed "s/^/base /"
    python benchmarks/one_gemm.py $shape | sed "s/^/new  /"
  done
done <<'S'
200704 1152 384 ln
200704 768 384 ln gelu
200704 384 384 res lo
200704 384 768 res lo
50176 2304 768 ln
50176 768 1536 res lo
173056 256 1024 relu
173056 256 2304 relu
200704 768 384
50176 2048 512 se
S
