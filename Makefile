# Builds the C-ABI library (sm_100a only), the native self-test and nothing else.
NVCC ?= nvcc
ARCH := -gencode arch=compute_100a,code=sm_100a
NVFLAGS := $(ARCH) -lineinfo -O3 -std=c++17 -Xcompiler -fPIC
PKG := vickers_hardness_unet_b200
SRC := $(wildcard $(PKG)/csrc/*.cuh) $(PKG)/csrc/capi.cu include/unetb200.h

all: $(PKG)/libunetb200.so build/selftest build/selftest_lib

$(PKG)/libunetb200.so: $(SRC)
	$(NVCC) $(NVFLAGS) -shared -o $@ $(PKG)/csrc/capi.cu

build/selftest: tests/native/selftest.cu tests/native/wconv2.cuh tests/native/wconv2_glue.cuh $(SRC)
	@mkdir -p build
	$(NVCC) $(NVFLAGS) -DUB_TC_PROF -Xcompiler -fopenmp -o $@ tests/native/selftest.cu

# the same self-test WITHOUT the role-cycle counters: exactly the kernel binaries libunetb200.so contains
build/selftest_lib: tests/native/selftest.cu tests/native/wconv2.cuh tests/native/wconv2_glue.cuh $(SRC)
	@mkdir -p build
	$(NVCC) $(NVFLAGS) -Xcompiler -fopenmp -o $@ tests/native/selftest.cu

clean:
	rm -f $(PKG)/libunetb200.so build/selftest build/selftest_lib
.PHONY: all clean
