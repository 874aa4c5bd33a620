# Build of libgiddy_b200 (product), the test oracle, and -- when /root/reference is present --
# the reference rebuilt for sm_100a (oracle/_ref, test infrastructure only).
#
#   make            product library + oracle
#   make ref        the reference oracle (needs /root/reference; see oracle/build_ref.sh)
#   make examples   C++ examples (link against the product library)
#   make sass       SASS / resource-usage listings into profiles/
#
# All device code is compiled for sm_100a only.

CUDA_HOME ?= /usr/local/cuda
NVCC      ?= $(CUDA_HOME)/bin/nvcc
# the environment's CXX may point at a wrapper without OpenMP support; use the system compiler
CXX       := /usr/bin/g++

ARCH_FLAGS := -gencode arch=compute_100a,code=sm_100a
NVCC_FLAGS := -std=c++17 -O3 -lineinfo $(ARCH_FLAGS) -Xcompiler -fPIC,-fvisibility=hidden -Iinclude -Iinclude/giddy --expt-relaxed-constexpr -diag-suppress 186,1160 $(EXTRA_DEFS)
CXX_FLAGS  := -std=c++17 -O2 -fPIC -fvisibility=hidden -Wall -Wextra -Wno-unused-parameter -Iinclude -Iinclude/giddy -I$(CUDA_HOME)/include $(EXTRA_DEFS)

SRC_DIR   := libgiddy_b200/csrc
BUILD_DIR ?= build
LIB_DIR   ?= libgiddy_b200/lib
LIB       := $(LIB_DIR)/libgiddy_b200.so

ABI_SRCS  := $(wildcard $(SRC_DIR)/abi/*.cu)
HOST_SRCS := $(wildcard $(SRC_DIR)/host/*.cpp)
ABI_OBJS  := $(patsubst $(SRC_DIR)/abi/%.cu,$(BUILD_DIR)/abi/%.o,$(ABI_SRCS))
HOST_OBJS := $(patsubst $(SRC_DIR)/host/%.cpp,$(BUILD_DIR)/host/%.o,$(HOST_SRCS))
KERNEL_HDRS := $(wildcard $(SRC_DIR)/kernels/*.cuh) $(wildcard $(SRC_DIR)/abi/*.hpp) include/giddy_b200.h $(wildcard include/giddy/*.h)
HOST_HDRS := $(wildcard include/giddy/*.hpp) $(wildcard include/giddy/*.h) $(wildcard include/giddy/*/*.hpp) $(wildcard include/giddy/*/*.h) include/giddy_b200.h

ORACLE_LIB := oracle/libgiddy_oracle.so
# The oracle doubles as the timed CPU baseline.  It is built in this container and travels to the GPU box as a binary, so it
# cannot be -march=native (the box's host CPU may differ: an illegal instruction there would take the tests down with it);
# x86-64-v3 (AVX2, BMI2, FMA) is the newest level both are sure to have.  bench.py rebuilds it -march=native on the box
# it is timed on when a compiler is there (oracle/libgiddy_oracle_native.so).
ORACLE_MARCH ?= x86-64-v3

.PHONY: all lib oracle ref examples sass clean variant

# A build of the library under other compile-time settings, for A/B measurements on the GPU box:
#   make variant NAME=p5 DEFS="-DGIDDY_PATCHED_BLOCKS_PER_MULTIPROCESSOR=5" REBUILD="abi/abi_patched host/registration"
#     -> benchmarks/variants/p5/libgiddy_b200.so
# Starts from the default build's objects and recompiles only the translation units named in REBUILD (those the settings
# touch; default: all).  Benchmarks load it with libgiddy_b200.Library(path); the .so is git-ignored but travels to the box.
REBUILD ?= $(patsubst $(BUILD_DIR)/%.o,%,$(ABI_OBJS) $(HOST_OBJS))
variant: lib
	@mkdir -p build/variants/$(NAME)/obj
	cp -r $(BUILD_DIR)/abi $(BUILD_DIR)/host build/variants/$(NAME)/obj/
	rm -f $(foreach unit,$(REBUILD),build/variants/$(NAME)/obj/$(unit).o)
	$(MAKE) -j8 lib BUILD_DIR=build/variants/$(NAME)/obj LIB_DIR=benchmarks/variants/$(NAME) EXTRA_DEFS="$(DEFS)"

all: lib oracle

lib: $(LIB)

$(BUILD_DIR)/abi/%.o: $(SRC_DIR)/abi/%.cu $(KERNEL_HDRS)
	@mkdir -p $(dir $@)
	$(NVCC) $(NVCC_FLAGS) -Xptxas -v -c $< -o $@ 2> $(BUILD_DIR)/abi/$*.ptxas.log || (cat $(BUILD_DIR)/abi/$*.ptxas.log; exit 1)

$(BUILD_DIR)/host/%.o: $(SRC_DIR)/host/%.cpp $(HOST_HDRS)
	@mkdir -p $(dir $@)
	$(CXX) $(CXX_FLAGS) -c $< -o $@

$(LIB): $(ABI_OBJS) $(HOST_OBJS)
	@mkdir -p $(LIB_DIR)
	$(NVCC) $(ARCH_FLAGS) -shared -o $@ $^ -cudart static -lpthread -ldl

oracle: $(ORACLE_LIB)

$(ORACLE_LIB): oracle/giddy_oracle.cpp
	$(CXX) -std=c++17 -O3 -march=$(ORACLE_MARCH) -fopenmp -fPIC -shared -Wall -o $@ $<

ref:
	bash oracle/build_ref.sh

examples: lib
	$(MAKE) -C examples

sass: lib
	@mkdir -p profiles/sass
	for o in $(ABI_OBJS); do b=$$(basename $$o .o); \
		$(CUDA_HOME)/bin/cuobjdump -sass $$o > profiles/sass/$$b.sass; \
		cp $(BUILD_DIR)/abi/$$b.ptxas.log profiles/sass/$$b.ptxas.txt; done

clean:
	rm -rf $(BUILD_DIR) $(LIB) $(ORACLE_LIB)
