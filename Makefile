# Plain build of the product library for users who do not drive it from Python (same flags as __graft_entry__.build()).
#   make            -> i-emic_b200/libthcm_b200.so   (nvcc, sm_100a; cross-compiles without a GPU)
#   make example    -> examples/newton_step          (C++ mirror include/thcm_model.hpp over the C ABI; needs a B200 to run)
#   make oracles    -> test infrastructure (CPU oracle, host emulation of the device functions, reference Krylov templates)
NVCC      ?= $(or $(CUDA_HOME),/usr/local/cuda)/bin/nvcc
NVCCFLAGS := -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 --fmad=false -std=c++17 --expt-relaxed-constexpr -Xcompiler -fPIC -shared
CSRC      := i-emic_b200/csrc
SRCS      := $(CSRC)/thcm_host.cpp $(CSRC)/thcm_probe.cpp $(CSRC)/thcm_assembly.cu $(CSRC)/thcm_linalg.cu $(CSRC)/thcm_api.cu
LIB       := i-emic_b200/libthcm_b200.so

all: $(LIB)

$(LIB): $(SRCS) $(wildcard $(CSRC)/*.h $(CSRC)/*.cuh) include/thcm_b200.h
	$(NVCC) $(NVCCFLAGS) -o $@ $(SRCS) -ldl

example: $(LIB) examples/newton_step.cpp include/thcm_model.hpp
	g++ -O2 -std=c++14 -Iinclude -o examples/newton_step examples/newton_step.cpp -Li-emic_b200 -lthcm_b200 -Wl,-rpath,'$$ORIGIN/../i-emic_b200' -ldl -lpthread

oracles:
	$(MAKE) -C oracle all
	$(MAKE) -C tests/cpp all

clean:
	rm -f $(LIB) examples/newton_step

.PHONY: all example oracles clean
