# Builds the product library (CUDA, sm_100a only), the C++ host layer / CLI and the CPU oracle.
NVCC      ?= /usr/local/cuda/bin/nvcc
HOSTCXX   := /usr/bin/g++
ARCH      := -gencode arch=compute_100a,code=sm_100a
NVFLAGS   := $(ARCH) -std=c++17 -O3 -lineinfo -ccbin $(HOSTCXX) -Xcompiler -fPIC,-fvisibility=hidden -I/usr/include
ifdef DEBUG_KERNELS
NVFLAGS   += -DVROD_KERNEL_DEBUG    # development stamps / timing modes (VROD_BATCHED_DEBUG, VROD_SCAN_DEBUG): never in production
endif
CSRC      := vrod_b200/csrc
OBJS      := $(CSRC)/knn_scan.o $(CSRC)/knn_batched.o $(CSRC)/vrod_capi.o
LIB       := vrod_b200/libvrod_knn.so

all: $(LIB) oracle host

$(CSRC)/%.o: $(CSRC)/%.cu $(wildcard $(CSRC)/*.cuh) include/vrod_knn.h
	$(NVCC) $(NVFLAGS) -c $< -o $@

$(LIB): $(OBJS)
	$(NVCC) $(ARCH) -shared -ccbin $(HOSTCXX) -o $@ $(OBJS) -ldl -lcuda

oracle:
	$(MAKE) -C oracle -s

host: $(LIB)
	@if [ -f vrod_b200/host/Makefile ]; then $(MAKE) -C vrod_b200/host -s; fi

clean:
	rm -f $(OBJS) $(LIB)
	$(MAKE) -C oracle clean

.PHONY: all oracle host clean
