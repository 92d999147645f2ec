# Builds the C-ABI shared library in-tree (the .so travels to the GPU box with the snapshot).
# One object per translation unit so that `make -j` compiles them side by side.
NVCC ?= nvcc
ARCH := -gencode arch=compute_100a,code=sm_100a
CSRC := phoskintime_b200/csrc
OBJD := build
LIB  := phoskintime_b200/libphoskin_b200.so
UNITS := pk_api pk_tps_dist pk_tps_succ pk_dense pk_global pk_nlls
OBJS := $(UNITS:%=$(OBJD)/%.o)
HDRS := $(wildcard $(CSRC)/*.cuh) $(CSRC)/pk_internal.hpp include/phoskin_b200.h
FLAGS := $(ARCH) -lineinfo -O3 -std=c++17 -Xptxas -v -Xcompiler -fPIC
MAKEFLAGS += -j8

$(LIB): $(OBJS)
	$(NVCC) $(ARCH) -shared -Xcompiler -fPIC -o $@ $(OBJS) -ldl
	@cat $(UNITS:%=$(OBJD)/%.ptxas.log) > $(CSRC)/ptxas.log
	@# zero local memory in EVERY kernel: no spill and no stack frame (a register array indexed at run time would show up here)
	@if grep -En '[1-9][0-9]* bytes (stack frame|spill stores|spill loads)' $(CSRC)/ptxas.log; then \
		echo "local memory in a kernel (see the lines above; $(CSRC)/ptxas.log)"; rm -f $@; exit 1; fi

$(OBJD)/%.o: $(CSRC)/%.cu $(HDRS)
	@mkdir -p $(OBJD)
	$(NVCC) $(FLAGS) -c -o $@ $< 2> $(OBJD)/$*.ptxas.log || (cat $(OBJD)/$*.ptxas.log; exit 1)

# debug build with per-phase cycle counters in the global-network kernel (tools/trace_global.py)
trace: $(UNITS:%=$(CSRC)/%.cu) $(HDRS)
	$(NVCC) $(ARCH) -lineinfo -O3 -std=c++17 -DPK_GLOBAL_TRACE -shared -Xcompiler -fPIC -o phoskintime_b200/libphoskin_b200_trace.so $(UNITS:%=$(CSRC)/%.cu) -ldl

clean:
	rm -rf $(LIB) $(CSRC)/ptxas.log $(OBJD)
.PHONY: clean trace
