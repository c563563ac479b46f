# Build of the B200-native NiftyMatch hot path.
#   make            -> niftymatch_b200/libnm_b200.so (C-ABI, sm_100a), oracle/libnm_oracle.so
#   make ref        -> oracle/_ref/libnmref.so (needs /root/reference; test infrastructure)
#   make compat     -> build/compat/{libgpuutils,libkernels,libsift}.a + include/nm layout
NVCC      ?= /usr/local/cuda/bin/nvcc
ARCH      := -gencode arch=compute_100a,code=sm_100a
NVFLAGS   := $(ARCH) -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xcompiler -Wall -Iinclude
CSRC      := niftymatch_b200/csrc
SRCS      := $(CSRC)/nm_pyramid.cu $(CSRC)/nm_extrema.cu $(CSRC)/nm_orient_desc.cu \
             $(CSRC)/nm_match.cu $(CSRC)/nm_match_tc.cu $(CSRC)/nm_sift.cu $(CSRC)/nm_ransac.cu $(CSRC)/nm_preprocess.cu $(CSRC)/nm_mosaic.cu
OBJS      := $(patsubst $(CSRC)/%.cu,build/obj/%.o,$(SRCS))
HDRS      := $(wildcard $(CSRC)/*.cuh) include/nm_b200.h
LIB       := niftymatch_b200/libnm_b200.so
MGPU      := niftymatch_b200/libnm_b200_mgpu.so

all: $(LIB) $(MGPU) oracle/libnm_oracle.so

build/obj/%.o: $(CSRC)/%.cu $(HDRS)
	@mkdir -p build/obj
	$(NVCC) $(NVFLAGS) -c $< -o $@

$(LIB): $(OBJS)
	$(NVCC) $(ARCH) -shared -o $@ $(OBJS)

# multi-GPU entries (include/nm_b200_mgpu.h): host orchestration over the C-ABI + NCCL
$(MGPU): $(CSRC)/nm_mgpu.cu include/nm_b200_mgpu.h $(LIB)
	$(NVCC) $(NVFLAGS) -shared $(CSRC)/nm_mgpu.cu -o $@ -Lniftymatch_b200 -lnm_b200 -lnccl -Xlinker -rpath -Xlinker '$$ORIGIN'

oracle/libnm_oracle.so: oracle/nm_oracle.c
	gcc -O2 -ffp-contract=off -fopenmp -shared -fPIC -Wall -o $@ $< -lm

ref:
	bash oracle/build_ref.sh

# Drop-in layer: the reference's header names, its three library names and its CMake package file,
# laid out like the reference's install tree (<prefix>/include/nm, <prefix>/lib/nm).
CPREFIX   := build/compat/prefix
CFLAGS_C  := $(ARCH) -O2 -std=c++17 -Xcompiler -fPIC -w -I$(CPREFIX)/include/nm
compat: $(LIB) $(MGPU)
	@mkdir -p $(CPREFIX)/include/nm $(CPREFIX)/lib/nm build/compat/obj
	cp compat/include/nm/*.h compat/include/nm/nm_compat.hpp include/nm_b200.h compat/NiftyMatchConfig.cmake $(CPREFIX)/include/nm/
	$(NVCC) $(CFLAGS_C) -c compat/src/compat_utils.cu -o build/compat/obj/compat_utils.o
	$(NVCC) $(CFLAGS_C) -c compat/src/compat_kernels.cu -o build/compat/obj/compat_kernels.o
	$(NVCC) $(CFLAGS_C) -c compat/src/compat_sift.cu -o build/compat/obj/compat_sift.o
	rm -f $(CPREFIX)/lib/nm/*.a
	ar rcs $(CPREFIX)/lib/nm/libgpuutils.a build/compat/obj/compat_utils.o
	ar rcs $(CPREFIX)/lib/nm/libkernels.a build/compat/obj/compat_kernels.o
	ar rcs $(CPREFIX)/lib/nm/libsift.a build/compat/obj/compat_sift.o
	cp $(LIB) $(CPREFIX)/lib/nm/libnm_b200.so
	cp $(MGPU) $(CPREFIX)/lib/nm/libnm_b200_mgpu.so
	cp include/nm_b200_mgpu.h $(CPREFIX)/include/nm/

# The client loop that drives the REFERENCE in the parity tests (oracle/ref_driver.cu), compiled
# unchanged against the drop-in tree: test artefact, entry points nmcompat_*.
compat-client: compat
	$(NVCC) $(CFLAGS_C) -DNM_COMPAT_BUILD -shared oracle/ref_driver.cu oracle/ref_ransac_driver.cu oracle/ref_preprocess_driver.cu oracle/ref_mosaic_driver.cu -o build/compat/libnmcompat.so \
	    -L$(CPREFIX)/lib/nm -lsift -lkernels -lgpuutils -lnm_b200 -Xlinker -rpath -Xlinker '$$ORIGIN/prefix/lib/nm'

clean:
	rm -rf build $(LIB) $(MGPU) oracle/libnm_oracle.so

.PHONY: all ref clean compat compat-client
