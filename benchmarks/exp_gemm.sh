for d in 0 16; do VIP_GEMM_DEBUG=$d python benchmarks/one_gemm.py 3211264 288 96 ln; done
for d in 0 16; do VIP_GEMM_DEBUG=$d python benchmarks/one_gemm.py 200704 1152 384 ln; done
python benchmarks/one_gemm.py 3211264 288 96
python benchmarks/one_gemm.py 200704 384 384
for d in 0 32; do VIP_GEMM_DEBUG=$d python benchmarks/one_gemm.py 200704 384 384 res; done
for d in 0 32 64 96; do VIP_GEMM_DEBUG=$d python benchmarks/one_gemm.py 200704 384 384 res lo; done
for d in 0 32 64 96; do VIP_GEMM_DEBUG=$d python benchmarks/one_gemm.py 3211264 96 96 res lo; done
python benchmarks/one_gemm.py 3211264 96 96 res
python benchmarks/one_gemm.py 3211264 96 96
