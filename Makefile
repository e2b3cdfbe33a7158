# Build everything in-tree (the .so files are git-ignored but travel to the GPU box).
NVCC      ?= nvcc
ARCH      := -gencode arch=compute_100a,code=sm_100a
NVFLAGS   := -O3 -std=c++17 -lineinfo $(ARCH) -Xcompiler -fPIC -Xptxas -v
PKG       := slam_maskrcnn_b200
LIB       := $(PKG)/libsfm_b200.so
CSRC      := $(PKG)/csrc/sfm_api.cu
CHDR      := $(wildcard $(PKG)/csrc/*.cuh) include/sfm_b200.h

all: $(LIB) oracle/liboracle.so driver/sfm_driver driver/sfm_driver_mgpu

$(LIB): $(CSRC) $(CHDR)
	$(NVCC) $(NVFLAGS) -shared -o $@ $(CSRC)

# CPU oracle (test infrastructure).  -ffp-contract=off: no FMA contraction beyond the explicit fmaf().
oracle/liboracle.so: oracle/sfm_oracle.c
	gcc -O2 -fPIC -shared -ffp-contract=off -mfma -fopenmp -o $@ $< -lm

driver/sfm_driver: driver/kernel.cpp driver/tum_io.hpp include/sfm_b200.hpp include/sfm_b200.h $(LIB)
	g++ -O2 -std=c++17 -pthread -Iinclude -o $@ driver/kernel.cpp -L$(PKG) -lsfm_b200 -Wl,-rpath,'$$ORIGIN/../$(PKG)' -lz

# the same driver over N GPUs of one box, NCCL called from C++ (system libnccl)
driver/sfm_driver_mgpu: driver/kernel_mgpu.cpp driver/tum_io.hpp include/sfm_b200.hpp include/sfm_b200.h $(LIB)
	g++ -O2 -std=c++17 -Iinclude -I/usr/local/cuda/include -o $@ driver/kernel_mgpu.cpp -L$(PKG) -lsfm_b200 -Wl,-rpath,'$$ORIGIN/../$(PKG)' \
		-L/usr/local/cuda/lib64 -lcudart -lnccl -lz

# A/B build of K1b with cp.async.bulk.tensor.2d depth tiles (measured slower, profiles/r2_k1_tma_tensor_ab.txt);
# select it with SFM_B200_LIB=$PWD/build/lib_tma.so
ab_tma:
	mkdir -p build && $(NVCC) $(NVFLAGS) -DSFM_K1_TMA_DEPTH=1 -shared -o build/lib_tma.so $(CSRC)

ref:
	python oracle/build_ref.py

clean:
	rm -f $(LIB) oracle/liboracle.so driver/sfm_driver driver/sfm_driver_mgpu

.PHONY: all ref clean ab_tma
