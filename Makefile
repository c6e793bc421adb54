# Builds the C-ABI library (sm_100a only), the native self-test and nothing else.
NVCC ?= nvcc
ARCH := -gencode arch=compute_100a,code=sm_100a
# -fno-gnu-unique: function-local statics of inline functions (the "max dynamic smem attribute set" flags of the launch
# helpers) must stay private to each of the two libraries; as STB_GNU_UNIQUE symbols they would be shared process-wide
NVFLAGS := $(ARCH) -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xcompiler -fno-gnu-unique
PKG := vickers_hardness_unet_b200
SRC := $(wildcard $(PKG)/csrc/*.cuh) $(PKG)/csrc/capi.cu include/unetb200.h

all: $(PKG)/libunetb200.so $(PKG)/libunetb200_f16.so build/selftest build/selftest_lib

# -Bsymbolic: the two libraries export the same C symbols and are loaded side by side (RTLD_LOCAL)
$(PKG)/libunetb200.so: $(SRC)
	$(NVCC) $(NVFLAGS) -shared -Xlinker -Bsymbolic -o $@ $(PKG)/csrc/capi.cu

# the same kernels with IEEE-half activations / operands instead of bfloat16 (inference only; see csrc/ptx.cuh UB_F16)
$(PKG)/libunetb200_f16.so: $(SRC)
	$(NVCC) $(NVFLAGS) -DUB_F16 -shared -Xlinker -Bsymbolic -o $@ $(PKG)/csrc/capi.cu

build/selftest: tests/native/selftest.cu tests/native/wconv2.cuh tests/native/wconv2_glue.cuh $(SRC)
	@mkdir -p build
	$(NVCC) $(NVFLAGS) -DUB_TC_PROF -Xcompiler -fopenmp -o $@ tests/native/selftest.cu

# the same self-test WITHOUT the role-cycle counters: exactly the kernel binaries libunetb200.so contains
build/selftest_lib: tests/native/selftest.cu tests/native/wconv2.cuh tests/native/wconv2_glue.cuh $(SRC)
	@mkdir -p build
	$(NVCC) $(NVFLAGS) -Xcompiler -fopenmp -o $@ tests/native/selftest.cu

clean:
	rm -f $(PKG)/libunetb200.so $(PKG)/libunetb200_f16.so build/selftest build/selftest_lib
.PHONY: all clean
