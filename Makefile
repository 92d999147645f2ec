# Builds the C-ABI shared library in-tree (the .so travels to the GPU box with the snapshot).
NVCC ?= nvcc
ARCH := -gencode arch=compute_100a,code=sm_100a
CSRC := phoskintime_b200/csrc
LIB  := phoskintime_b200/libphoskin_b200.so
SRCS := $(CSRC)/pk_api.cu $(CSRC)/pk_global.cu
HDRS := $(wildcard $(CSRC)/*.cuh) $(CSRC)/pk_internal.hpp include/phoskin_b200.h

$(LIB): $(SRCS) $(HDRS)
	$(NVCC) $(ARCH) -lineinfo -O3 -std=c++17 -Xptxas -v -shared -Xcompiler -fPIC -o $@ $(SRCS) -ldl 2> $(CSRC)/ptxas.log || (cat $(CSRC)/ptxas.log; exit 1)
	@grep -c "0 bytes spill stores" $(CSRC)/ptxas.log >/dev/null

# debug build with per-phase cycle counters in the global-network kernel (tools/trace_global.py)
trace: $(SRCS) $(HDRS)
	$(NVCC) $(ARCH) -lineinfo -O3 -std=c++17 -DPK_GLOBAL_TRACE -shared -Xcompiler -fPIC -o phoskintime_b200/libphoskin_b200_trace.so $(SRCS) -ldl

clean:
	rm -f $(LIB) $(CSRC)/ptxas.log
.PHONY: clean trace
