# Builds the product library (CUDA, sm_100a only), the C++ host layer / CLI and the CPU oracle.
NVCC      ?= /usr/local/cuda/bin/nvcc
HOSTCXX   := /usr/bin/g++
ARCH      := -gencode arch=compute_100a,code=sm_100a
NVFLAGS   := $(ARCH) -std=c++17 -O3 -lineinfo -ccbin $(HOSTCXX) -Xcompiler -fPIC,-fvisibility=hidden -I/usr/include
CSRC      := vrod_b200/csrc
# `make debug` builds vrod_b200/libvrod_knn_dbg.so with the development stamps / timing modes compiled in
# (VROD_BATCHED_DEBUG, VROD_SCAN_DEBUG); VROD_LIB=<path> makes vrod_b200/ffi.py load it.  Never the production library.
ifdef DEBUG_KERNELS
NVFLAGS   += -DVROD_KERNEL_DEBUG
SUF       := _dbg
endif
OBJS      := $(CSRC)/knn_scan$(SUF).o $(CSRC)/knn_batched$(SUF).o $(CSRC)/vrod_capi$(SUF).o
LIB       := vrod_b200/libvrod_knn$(SUF).so

all: $(LIB) oracle host

$(CSRC)/%$(SUF).o: $(CSRC)/%.cu $(wildcard $(CSRC)/*.cuh) $(wildcard $(CSRC)/*.hpp) include/vrod_knn.h
	$(NVCC) $(NVFLAGS) -c $< -o $@

debug:
	$(MAKE) DEBUG_KERNELS=1 vrod_b200/libvrod_knn_dbg.so

$(LIB): $(OBJS)
	$(NVCC) $(ARCH) -shared -ccbin $(HOSTCXX) -o $@ $(OBJS) -ldl -lcuda

oracle:
	$(MAKE) -C oracle -s

host: $(LIB)
	@if [ -f vrod_b200/host/Makefile ]; then $(MAKE) -C vrod_b200/host -s; fi

clean:
	rm -f $(CSRC)/*.o vrod_b200/libvrod_knn.so vrod_b200/libvrod_knn_dbg.so
	$(MAKE) -C oracle clean

.PHONY: all oracle host clean debug
