# Top-level build for C users of the backend (the Python tests/bench use ugemm_b200/build.py, same flags).
NVCC  ?= nvcc
ARCH  := -gencode arch=compute_100a,code=sm_100a
SRC   := ugemm_b200/csrc/backend.cu ugemm_b200/csrc/k1_tcgen05.cu ugemm_b200/csrc/k2_simt.cu ugemm_b200/csrc/k3_level12.cu ugemm_b200/csrc/k4_dgemm.cu ugemm_b200/csrc/shard.cu
LIB   := ugemm_b200/libugemm_cuda.so

all: $(LIB) oracle harness
$(LIB): $(SRC) ugemm_b200/csrc/common.cuh ugemm_b200/csrc/ptx.cuh ugemm_b200/csrc/k1_common.cuh ugemm_b200/csrc/k1_ss.cuh ugemm_b200/csrc/k1_ts.cuh include/ugemm_cuda.h
	$(NVCC) $(ARCH) -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -shared -o $@ $(SRC) -ldl
oracle:
	$(MAKE) -C oracle all
harness: $(LIB)
	$(MAKE) -C harness all
check: all
	cd harness && ./check_sgemm_cuda M=1024 N=1024 K=1024 && ./sgemm_test_cuda
# compute-sanitizer over small cases of every kernel (K1 TS and SS, stream-K tail, K2, fused convolution, level 1/2, DGEMM).
# racecheck is expected to report ONE hazard class only: the shared-memory word that tcgen05.alloc writes the TMEM base address
# into -- the write is made by the allocation hardware and ordered for its readers by tcgen05.fence + the CTA / cluster barrier,
# which racecheck does not model.  Needs a B200; the log goes to profiles/.
# (SANITIZE_CASES=tools/sanitize_beta.py SANITIZE_TOOLS="memcheck racecheck": the beta != 0 path alone)
SANITIZE_CASES ?= tools/sanitize_cases.py
SANITIZE_TOOLS ?= memcheck synccheck racecheck
sanitize: $(LIB) oracle
	@mkdir -p gpurun_out
	@for tool in $(SANITIZE_TOOLS); do \
	  echo "== compute-sanitizer --tool $$tool"; \
	  compute-sanitizer --tool $$tool --print-limit 20 python $(SANITIZE_CASES) > gpurun_out/sanitize_$$tool.log 2>&1; echo "rc=$$?"; \
	  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|hazard|Error|relerr|conv|ok" gpurun_out/sanitize_$$tool.log | sort | uniq -c | sort -rn | head -40; \
	done
clean:
	rm -f $(LIB); $(MAKE) -C harness clean; $(MAKE) -C oracle clean
.PHONY: all oracle harness check sanitize clean
