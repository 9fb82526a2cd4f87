# Builds libturbosqueeze_b200.so (the product: CUDA kernels + C-ABI, sm_100a only) and the
# workload generator.  `python -c "import __graft_entry__ as g; g.build()"` runs this and oracle/Makefile.
NVCC   ?= nvcc
CC     ?= gcc
PKG    := turbosqueeze_b200
CSRC   := $(PKG)/csrc
ARCH   := -gencode arch=compute_100a,code=sm_100a
NVFLAGS := $(ARCH) -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xcompiler -Wall -Xcompiler -Wno-unused-function
CU     := $(CSRC)/tsq_decode_split.cu $(CSRC)/tsq_encode_scalar.cu $(CSRC)/tsq_encode_batch.cu $(CSRC)/tsq_container.cu $(CSRC)/tsq_capi.cu
OBJ    := $(CU:.cu=.o)
HDR    := $(wildcard $(CSRC)/*.cuh) include/tsq_b200.h
# Test-only cross-check library: the product's objects + the superseded round-1 kernels (csrc/xcheck/), reachable
# through encode_impl = 2 / decode_lanes = 1..33.  Never loaded by bench.py or the product path.
XCU    := $(CSRC)/xcheck/tsq_decode_subwarp.cu $(CSRC)/xcheck/tsq_decode_warp.cu $(CSRC)/xcheck/tsq_encode_warp.cu
XOBJ   := $(XCU:.cu=.o) $(CSRC)/xcheck/tsq_container_x.o
XLIB   := tests/xcheck/libturbosqueeze_b200_xcheck.so

all: $(PKG)/libturbosqueeze_b200.so $(PKG)/libtsq_workload.so $(XLIB)

$(CSRC)/xcheck/tsq_container_x.o: $(CSRC)/tsq_container.cu $(HDR)
	$(NVCC) $(NVFLAGS) -DTSQB_XCHECK -c $< -o $@

$(CSRC)/xcheck/%.o: $(CSRC)/xcheck/%.cu $(HDR)
	$(NVCC) $(NVFLAGS) -c $< -o $@

$(XLIB): $(XOBJ) $(filter-out $(CSRC)/tsq_container.o,$(OBJ))
	mkdir -p tests/xcheck
	$(NVCC) $(ARCH) -shared -o $@ $^ -lpthread

$(CSRC)/%.o: $(CSRC)/%.cu $(HDR)
	$(NVCC) $(NVFLAGS) -c $< -o $@

$(PKG)/libturbosqueeze_b200.so: $(OBJ)
	$(NVCC) $(ARCH) -shared -o $@ $(OBJ) -lpthread

$(PKG)/libtsq_workload.so: $(CSRC)/tsq_workload.c
	$(CC) -O2 -fPIC -shared -o $@ $< -lm -lpthread

clean:
	rm -f $(OBJ) $(XOBJ) $(XLIB) $(PKG)/libturbosqueeze_b200.so $(PKG)/libtsq_workload.so
.PHONY: all clean
