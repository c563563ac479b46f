# Build of the B200-native NiftyMatch hot path.
#   make            -> niftymatch_b200/libnm_b200.so (C-ABI, sm_100a), oracle/libnm_oracle.so
#   make ref        -> oracle/_ref/libnmref.so (needs /root/reference; test infrastructure)
#   make compat     -> build/compat/{libgpuutils,libkernels,libsift}.a + include/nm layout
NVCC      ?= /usr/local/cuda/bin/nvcc
ARCH      := -gencode arch=compute_100a,code=sm_100a
NVFLAGS   := $(ARCH) -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xcompiler -Wall -Iinclude
CSRC      := niftymatch_b200/csrc
SRCS      := $(CSRC)/nm_pyramid.cu $(CSRC)/nm_extrema.cu $(CSRC)/nm_orient_desc.cu \
             $(CSRC)/nm_match.cu $(CSRC)/nm_match_tc.cu $(CSRC)/nm_sift.cu
OBJS      := $(patsubst $(CSRC)/%.cu,build/obj/%.o,$(SRCS))
HDRS      := $(wildcard $(CSRC)/*.cuh) include/nm_b200.h
LIB       := niftymatch_b200/libnm_b200.so

all: $(LIB) oracle/libnm_oracle.so

build/obj/%.o: $(CSRC)/%.cu $(HDRS)
	@mkdir -p build/obj
	$(NVCC) $(NVFLAGS) -c $< -o $@

$(LIB): $(OBJS)
	$(NVCC) $(ARCH) -shared -o $@ $(OBJS)

oracle/libnm_oracle.so: oracle/nm_oracle.c
	gcc -O2 -ffp-contract=off -fopenmp -shared -fPIC -Wall -o $@ $< -lm

ref:
	bash oracle/build_ref.sh

clean:
	rm -rf build $(LIB) oracle/libnm_oracle.so

.PHONY: all ref clean
