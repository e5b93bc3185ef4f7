# Builds libumgap_gpu.so (CUDA kernels + C ABI, sm_100a only), the `umgap` CLI on top of it, and
# the CPU oracle's C restatement.  Artefacts stay in-tree (git-ignored) so they travel to the GPU box.
NVCC      ?= /usr/local/cuda/bin/nvcc
CXX       ?= g++
CC        ?= gcc
ARCH      := -gencode arch=compute_100a,code=sm_100a
NVFLAGS   := $(ARCH) -O3 -std=c++17 -lineinfo -Xcompiler -fPIC,-Wall,-Wno-unused-function --expt-relaxed-constexpr $(EXTRA_NVFLAGS)
CSRC      := umgap_b200/csrc
LIBDIR    := umgap_b200/lib
BINDIR    := umgap_b200/bin
CU_SRCS   := $(wildcard $(CSRC)/*.cu)
CPP_SRCS  := $(wildcard $(CSRC)/*.cpp)
OBJS      := $(patsubst $(CSRC)/%.cu,build/%.o,$(CU_SRCS)) $(patsubst $(CSRC)/%.cpp,build/%.cpp.o,$(CPP_SRCS))
HDRS      := $(wildcard $(CSRC)/*.h) $(wildcard $(CSRC)/*.cuh) include/umgap_gpu.h

all: $(LIBDIR)/libumgap_gpu.so cli oracle

$(LIBDIR)/libumgap_gpu.so: $(OBJS)
	@mkdir -p $(LIBDIR)
	$(NVCC) $(ARCH) -shared -o $@ $(OBJS) -cudart shared

build/%.o: $(CSRC)/%.cu $(HDRS)
	@mkdir -p build
	$(NVCC) $(NVFLAGS) -c $< -o $@

build/%.cpp.o: $(CSRC)/%.cpp $(HDRS)
	@mkdir -p build
	$(CXX) -O2 -std=c++17 -fPIC -Wall -I/usr/local/cuda/include -c $< -o $@

cli: $(BINDIR)/umgap
$(BINDIR)/umgap: $(wildcard $(CSRC)/cli/*.cpp) $(LIBDIR)/libumgap_gpu.so include/umgap_gpu.h
	@mkdir -p $(BINDIR)
	@if ls $(CSRC)/cli/*.cpp >/dev/null 2>&1; then \
	  $(CXX) -O2 -std=c++17 -Wall -o $@ $(CSRC)/cli/*.cpp -Iinclude -L$(LIBDIR) -lumgap_gpu -Wl,-rpath,'$$ORIGIN/../lib' -lpthread; \
	fi

oracle:
	@if [ -f oracle/c/Makefile ]; then $(MAKE) -C oracle/c; fi

clean:
	rm -rf build $(LIBDIR)/*.so $(BINDIR)/umgap oracle/c/*.so

.PHONY: all cli oracle clean
