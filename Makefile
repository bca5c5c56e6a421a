# Builds libb4r.so (sm_100a only) in-tree and the oracle helpers.  `make -j` ; `make clean`.
NVCC ?= nvcc
ARCH := -gencode arch=compute_100a,code=sm_100a
NVFLAGS := $(ARCH) -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xptxas -v $(EXTRA)
SRC_DIR := bert4rec_b200/csrc
SRCS := $(wildcard $(SRC_DIR)/*.cu)
OBJS := $(patsubst $(SRC_DIR)/%.cu,build/%.o,$(SRCS))
LIB := bert4rec_b200/libb4r.so

all: $(LIB)

build/%.o: $(SRC_DIR)/%.cu $(wildcard $(SRC_DIR)/*.cuh) $(SRC_DIR)/kernels.h include/b4r.h include/b4r_debug.h
	@mkdir -p build
	$(NVCC) $(NVFLAGS) -c $< -o $@ 2> build/$*.ptxas.log || (cat build/$*.ptxas.log; exit 1)

$(LIB): $(OBJS)
	$(NVCC) $(ARCH) -shared -o $@ $(OBJS) -lcudart

clean:
	rm -rf build $(LIB)
.PHONY: all clean
