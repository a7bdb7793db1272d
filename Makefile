# Builds libacan_b200.so (sm_100a only) and nothing else.  `python __graft_entry__.py build` calls this.
NVCC      ?= /usr/local/cuda/bin/nvcc
ARCH      := -gencode arch=compute_100a,code=sm_100a
NVCCFLAGS := $(ARCH) -O3 -std=c++17 -lineinfo -Xcompiler -fPIC,-Wall,-Wno-unused-function --expt-relaxed-constexpr -Xptxas -v
SRC       := $(wildcard acan_b200/csrc/*.cu)
OBJ       := $(patsubst acan_b200/csrc/%.cu,build/%.o,$(SRC))
LIB       := acan_b200/lib/libacan_b200.so

all: $(LIB)

build/%.o: acan_b200/csrc/%.cu acan_b200/csrc/common.cuh include/acan_b200.h
	@mkdir -p build
	$(NVCC) $(NVCCFLAGS) -c $< -o $@ 2> build/$*.ptxas.log || (cat build/$*.ptxas.log; false)

$(LIB): $(OBJ)
	@mkdir -p acan_b200/lib
	$(NVCC) $(ARCH) -shared -o $@ $(OBJ) -cudart static

clean:
	rm -rf build acan_b200/lib

.PHONY: all clean
