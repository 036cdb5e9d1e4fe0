# Builds the C-ABI shared library for sm_100a (no torch dependency).
NVCC ?= nvcc
ARCH := -gencode arch=compute_100a,code=sm_100a
NVFLAGS := $(ARCH) -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xptxas -v
SRC := distbo_b200/csrc/chol.cu distbo_b200/csrc/gram.cu distbo_b200/csrc/score.cu distbo_b200/csrc/embed.cu distbo_b200/csrc/inverse_i8.cu distbo_b200/csrc/blr.cu distbo_b200/csrc/hostutil.cu
OBJ := $(SRC:.cu=.o)
LIB := distbo_b200/libdistbo_b200.so

all: $(LIB)

%.o: %.cu distbo_b200/csrc/common.cuh distbo_b200/csrc/gemm_core.cuh distbo_b200/csrc/potf2.cuh distbo_b200/csrc/tma.cuh distbo_b200/csrc/digit_planes.cuh include/distbo_b200.h
	$(NVCC) $(NVFLAGS) -c $< -o $@

$(LIB): $(OBJ)
	$(NVCC) $(ARCH) -shared -o $@ $(OBJ) -lcudart

clean:
	rm -f $(OBJ) $(LIB)
