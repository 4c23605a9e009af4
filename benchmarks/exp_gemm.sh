# epilogue cost experiments on the GCViT shapes at batch 1024 (VIP_GEMM_DEBUG bits: see GemmArgs.dbg in csrc/gemm.cu)
python benchmarks/one_gemm.py 200704 384 384
python benchmarks/one_gemm.py 200704 384 384 res
for d in 0 64; do VIP_GEMM_DEBUG=$d python benchmarks/one_gemm.py 200704 384 384 res lo; done
for d in 0 64; do VIP_GEMM_DEBUG=$d python benchmarks/one_gemm.py 200704 384 768 res lo; done
python benchmarks/one_gemm.py 200704 1152 384 ln
python benchmarks/one_gemm.py 200704 768 384 ln gelu
for d in 0 64; do VIP_GEMM_DEBUG=$d python benchmarks/one_gemm.py 802816 192 192 res lo; done
for d in 0 64; do VIP_GEMM_DEBUG=$d python benchmarks/one_gemm.py 3211264 96 96 res lo; done
python benchmarks/one_gemm.py 173056 1024 256 res
python benchmarks/one_gemm.py 2560000 256 64 res
