# Builds the product library (sm_100a only), the host-emulation test harness and the C++ CPU oracle.
#   make lib       -> mathlib_b200/libb200math.so      (nvcc, ships to the GPU box in-tree)
#   make hostemu   -> tests/hostemu/libhostemu.so      (g++; device headers compiled for the CPU, tests only)
#   make oracle    -> oracle/cpu/liboracle_cpu.so      (g++; CPU restatement used by tests and bench baselines)
NVCC ?= nvcc
CXX ?= g++
ARCH := -gencode arch=compute_100a,code=sm_100a
NVFLAGS := $(ARCH) -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xptxas -v
CSRC := mathlib_b200/csrc
HDRS := $(wildcard $(CSRC)/*.cuh) $(CSRC)/constants.h $(CSRC)/microcode.h include/b200.h
OBJS := $(CSRC)/build/abi.o $(CSRC)/build/kernels_bn254.o $(CSRC)/build/kernels_bls381.o $(CSRC)/build/kernels_bls377.o

.PHONY: all lib hostemu oracle clean
all: lib hostemu oracle

lib: mathlib_b200/libb200math.so

$(CSRC)/build/%.o: $(CSRC)/%.cu $(HDRS)
	@mkdir -p $(CSRC)/build
	$(NVCC) $(NVFLAGS) -c $< -o $@ 2> $(CSRC)/build/$*.ptxas.log || (cat $(CSRC)/build/$*.ptxas.log; exit 1)

mathlib_b200/libb200math.so: $(OBJS)
	$(NVCC) $(ARCH) -shared -o $@ $(OBJS) -lcudart

hostemu: tests/hostemu/libhostemu.so
tests/hostemu/libhostemu.so: tests/hostemu/hostemu.cpp $(HDRS)
	$(CXX) -O2 -std=c++17 -shared -fPIC -o $@ tests/hostemu/hostemu.cpp

oracle: oracle/cpu/liboracle_cpu.so
oracle/cpu/liboracle_cpu.so: $(wildcard oracle/cpu/*.cpp oracle/cpu/*.h)
	$(CXX) -O3 -march=native -std=c++17 -shared -fPIC -pthread -o $@ oracle/cpu/oracle_cpu.cpp

clean:
	rm -rf $(CSRC)/build mathlib_b200/libb200math.so tests/hostemu/libhostemu.so oracle/cpu/liboracle_cpu.so
