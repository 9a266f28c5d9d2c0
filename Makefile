# Top-level build for C users of the backend (the Python tests/bench use ugemm_b200/build.py, same flags).
NVCC  ?= nvcc
ARCH  := -gencode arch=compute_100a,code=sm_100a
SRC   := ugemm_b200/csrc/backend.cu ugemm_b200/csrc/k1_tcgen05.cu ugemm_b200/csrc/k2_simt.cu ugemm_b200/csrc/k3_level12.cu ugemm_b200/csrc/k4_dgemm.cu ugemm_b200/csrc/shard.cu
LIB   := ugemm_b200/libugemm_cuda.so

all: $(LIB) oracle harness
$(LIB): $(SRC) ugemm_b200/csrc/common.cuh ugemm_b200/csrc/ptx.cuh include/ugemm_cuda.h
	$(NVCC) $(ARCH) -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -shared -o $@ $(SRC) -ldl
oracle:
	$(MAKE) -C oracle all
harness: $(LIB)
	$(MAKE) -C harness all
check: all
	cd harness && ./check_sgemm_cuda M=1024 N=1024 K=1024 && ./sgemm_test_cuda
clean:
	rm -f $(LIB); $(MAKE) -C harness clean; $(MAKE) -C oracle clean
.PHONY: all oracle harness check clean
