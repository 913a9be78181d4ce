# Build of the B200 library (sm_100a only).  Outputs stay in-tree under prealps_b200/lib and
# prealps_b200/bin (git-ignored, shipped to the GPU box by gpurun).
#   make            libraries + the unchanged reference driver (when /root/reference is present)
#   make oracle     the test oracle (oracle/Makefile; needs /root/reference)
NVCC     ?= nvcc
CC       ?= gcc
CXX      ?= g++
REF      ?= /root/reference
METIS_A  ?= /usr/local/cuda/targets/x86_64-linux/lib/libmetis_static.a
ARCH     := -gencode arch=compute_100a,code=sm_100a
NVFLAGS  := $(ARCH) -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xcompiler -Wall
CFLAGS   := -O2 -fPIC -std=gnu99 -Wall -Wno-unused-result
LIB      := prealps_b200/lib
BIN      := prealps_b200/bin
OBJ      := build/obj

CU_SRCS  := ctx.cu spmm.cu ecg_kernels.cu bj_factor.cu bj_solve.cu
CU_OBJS  := $(patsubst %.cu,$(OBJ)/%.o,$(CU_SRCS)) $(OBJ)/bj_symbolic.o
H_SRCS   := pa_csr.c pa_operator.c pa_block_jacobi.c pa_ecg.c pa_driver.c
H_OBJS   := $(patsubst %.c,$(OBJ)/%.o,$(H_SRCS))

TARGETS := $(LIB)/libmpishim.so $(LIB)/libprealps_cuda.so $(LIB)/libprealps_b200.so $(BIN)/bench_kernels
ifneq ($(wildcard $(REF)/examples/test_ecg_prealps_op.c),)
TARGETS += $(BIN)/test_ecg_prealps_op $(BIN)/test_ecg_bench_fused
endif

all: $(TARGETS)

$(OBJ) $(LIB) $(BIN):
	mkdir -p $@

$(OBJ)/%.o: prealps_b200/csrc/%.cu prealps_b200/csrc/common.cuh prealps_b200/csrc/bj.h prealps_b200/csrc/bj_symbolic.h prealps_b200/csrc/spmm_kernels.cuh include/prealps_cuda.h | $(OBJ)
	$(NVCC) $(NVFLAGS) -c $< -o $@

$(OBJ)/bj_symbolic.o: prealps_b200/csrc/bj_symbolic.cpp prealps_b200/csrc/bj_symbolic.h | $(OBJ)
	$(CXX) -O2 -fPIC -std=c++17 -Wall -c $< -o $@

$(OBJ)/%.o: prealps_b200/host/%.c prealps_b200/host/pa_internal.h $(wildcard include/*.h) | $(OBJ)
	$(CC) $(CFLAGS) -Iinclude -Impishim -c $< -o $@

$(LIB)/libmpishim.so: mpishim/mpishim.c mpishim/mpi.h | $(LIB)
	$(CC) $(CFLAGS) -shared -Impishim $< -o $@ -lpthread

$(LIB)/libprealps_cuda.so: $(CU_OBJS) | $(LIB)
	$(NVCC) $(ARCH) -shared -o $@ $(CU_OBJS) $(METIS_A) -Xlinker --exclude-libs=ALL -ldl -lpthread

$(LIB)/libprealps_b200.so: $(H_OBJS) $(LIB)/libprealps_cuda.so $(LIB)/libmpishim.so | $(LIB)
	$(CC) -shared -o $@ $(H_OBJS) $(METIS_A) -Wl,--exclude-libs=ALL -L$(LIB) -lprealps_cuda -lmpishim \
	    -Wl,-rpath,'$$ORIGIN' -lm -lpthread

# the reference driver, compiled UNCHANGED straight from the reference tree
$(BIN)/test_ecg_prealps_op: $(REF)/examples/test_ecg_prealps_op.c $(LIB)/libprealps_b200.so | $(BIN)
	$(CC) -O2 -std=gnu99 -w -Iinclude/compat -Iinclude -Impishim $< -o $@ -L$(LIB) -lprealps_b200 -lprealps_cuda \
	    -lmpishim -Wl,-rpath,'$$ORIGIN/../lib' -lm

$(BIN)/test_ecg_bench_fused: $(REF)/examples/test_ecg_bench_fused.c $(LIB)/libprealps_b200.so | $(BIN)
	$(CC) -O2 -std=gnu99 -w -Iinclude/compat -Iinclude -Impishim $< -o $@ -L$(LIB) -lprealps_b200 -lprealps_cuda \
	    -lmpishim -Wl,-rpath,'$$ORIGIN/../lib' -lm

# the preAlps half of the reference's test_bench_spmm.c / test_bench_bjacobi.c (those need PETSc)
$(BIN)/bench_kernels: examples/bench_kernels.c $(LIB)/libprealps_b200.so | $(BIN)
	$(CC) -O2 -std=gnu99 -Wall -Iinclude -Impishim $< -o $@ -L$(LIB) -lprealps_b200 -lprealps_cuda \
	    -lmpishim -Wl,-rpath,'$$ORIGIN/../lib' -lm

oracle:
	$(MAKE) -C oracle

clean:
	rm -rf build $(LIB) $(BIN)
.PHONY: all oracle clean
