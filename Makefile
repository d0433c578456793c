# Convenience wrapper around sigfish_b200/build.py (same commands): `make -j` builds libsfgpu.so (nvcc, sm_100a only),
# libsfhost.so and the sigfish-b200 binary in sigfish_b200/; `make oracle` builds the test oracle.
NVCC     ?= /usr/local/cuda/bin/nvcc
CC       ?= gcc
PKG      := sigfish_b200
NVFLAGS  := -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -fmad=false -prec-div=true -prec-sqrt=true \
            -Xcompiler -fPIC
CFLAGS   := -O2 -g -std=c99 -Wall -D_GNU_SOURCE -Iinclude
HOST_SRC := $(filter-out $(PKG)/host/main.c,$(wildcard $(PKG)/host/*.c))
GPU_OBJ  := $(patsubst $(PKG)/csrc/%.cu,build/obj/%.o,$(wildcard $(PKG)/csrc/*.cu))

.PHONY: all oracle test clean
all: $(PKG)/sigfish-b200

# sfgpu.cu = the C-ABI and most kernels; sf_inst_*.cu = the instantiations of the pair DTW kernel (compiled side by side)
build/obj/%.o: $(PKG)/csrc/%.cu $(wildcard $(PKG)/csrc/*.cuh) include/sfgpu.h
	@mkdir -p build/obj
	$(NVCC) $(NVFLAGS) -c -o $@ $<

$(PKG)/libsfgpu.so: $(GPU_OBJ)
	$(NVCC) -gencode arch=compute_100a,code=sm_100a -shared -o $@ $(GPU_OBJ)

$(PKG)/libsfhost.so: $(HOST_SRC) $(wildcard $(PKG)/host/*.h) include/sfgpu.h $(PKG)/libsfgpu.so
	$(CC) $(CFLAGS) -fPIC -shared -o $@ $(HOST_SRC) -L$(PKG) -lsfgpu -Wl,-rpath,'$$ORIGIN' -lz -lpthread -lm

$(PKG)/sigfish-b200: $(PKG)/host/main.c $(PKG)/libsfhost.so
	$(CC) $(CFLAGS) -o $@ $< -L$(PKG) -lsfhost -lsfgpu -Wl,-rpath,'$$ORIGIN' -lz -lpthread -lm

oracle:
	$(MAKE) -C oracle all

test: all oracle
	python -m pytest tests -q -m "not gpu"

clean:
	rm -rf build/obj $(PKG)/libsfgpu.so $(PKG)/libsfhost.so $(PKG)/sigfish-b200
