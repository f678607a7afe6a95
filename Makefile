# Builds the B200 backend shared library (C ABI in include/s2c_b200.h) for sm_100a, in-tree.
NVCC      ?= nvcc
ARCH      := -gencode arch=compute_100a,code=sm_100a
NVCCFLAGS := $(ARCH) -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xcompiler -Wall
CSRC      := zk_symmetric_crypto_b200/csrc
SRCS      := $(wildcard $(CSRC)/*.cu)
CXXSRCS   := $(wildcard $(CSRC)/*.cpp)
OBJS      := $(patsubst $(CSRC)/%.cu,build/%.o,$(SRCS)) $(patsubst $(CSRC)/%.cpp,build/%.o,$(CXXSRCS))
HDRS      := $(wildcard $(CSRC)/*.cuh) $(wildcard $(CSRC)/*.hpp) include/s2c_b200.h
LIB       := zk_symmetric_crypto_b200/libs2c_b200.so

all: $(LIB)

build/%.o: $(CSRC)/%.cu $(HDRS)
	@mkdir -p build
	$(NVCC) $(NVCCFLAGS) -c $< -o $@

# host-only translation units (vector intrinsics with per-function targets and run-time dispatch; no -march flag)
build/%.o: $(CSRC)/%.cpp $(HDRS)
	@mkdir -p build
	$(CXX) -O3 -std=c++17 -fPIC -Wall -c $< -o $@

$(LIB): $(OBJS)
	$(NVCC) $(ARCH) -shared -o $@ $(OBJS) -lcudart -ldl

oracle-ref:
	$(MAKE) -C oracle ref

clean:
	rm -rf build $(LIB)
.PHONY: all clean oracle-ref
