# Builds libturbosqueeze_b200.so (the product: CUDA kernels + C-ABI, sm_100a only) and the
# workload generator.  `python -c "import __graft_entry__ as g; g.build()"` runs this and oracle/Makefile.
NVCC   ?= nvcc
CC     ?= gcc
PKG    := turbosqueeze_b200
CSRC   := $(PKG)/csrc
ARCH   := -gencode arch=compute_100a,code=sm_100a
NVFLAGS := $(ARCH) -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xcompiler -Wall -Xcompiler -Wno-unused-function
CU     := $(CSRC)/tsq_decode.cu $(CSRC)/tsq_decode_warp.cu $(CSRC)/tsq_decode_split.cu $(CSRC)/tsq_encode_scalar.cu $(CSRC)/tsq_encode_warp.cu $(CSRC)/tsq_encode_batch.cu $(CSRC)/tsq_container.cu $(CSRC)/tsq_capi.cu
OBJ    := $(CU:.cu=.o)
HDR    := $(wildcard $(CSRC)/*.cuh) include/tsq_b200.h

all: $(PKG)/libturbosqueeze_b200.so $(PKG)/libtsq_workload.so

$(CSRC)/%.o: $(CSRC)/%.cu $(HDR)
	$(NVCC) $(NVFLAGS) -c $< -o $@

$(PKG)/libturbosqueeze_b200.so: $(OBJ)
	$(NVCC) $(ARCH) -shared -o $@ $(OBJ) -lpthread

$(PKG)/libtsq_workload.so: $(CSRC)/tsq_workload.c
	$(CC) -O2 -fPIC -shared -o $@ $< -lm -lpthread

clean:
	rm -f $(OBJ) $(PKG)/libturbosqueeze_b200.so $(PKG)/libtsq_workload.so
.PHONY: all clean
