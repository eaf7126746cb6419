# Build of the B200 D2Q9-BGK solver.
#
#   make            liblbm_b200.so (sm_100a CUDA + C ABI) and the d2q9-bgk host program
#   make oracle     the CPU checker under oracle/ (test infrastructure; never linked into the product)
#   make check      run d2q9-bgk on the shipped 128x128 case and validate its files with the reference's own
#                   check.py (tests/ref_check/check.py, byte for byte; SerialCode/Makefile:20-25).  Needs a GPU.
#   make check-all  the same for all four shipped grids at their full iteration counts (LBM_ARITH=fast to
#                   check the fast flavour)
#   make clean
#
# nvcc cross-compiles for sm_100a without a GPU.

NVCC      ?= /usr/local/cuda/bin/nvcc
HOSTCC    := /usr/bin/gcc
PKG       := lbm-asynchronous_b200
CUDA_ARCH := -gencode arch=compute_100a,code=sm_100a
NVFLAGS   := $(CUDA_ARCH) -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xptxas -v
LIB       := $(PKG)/liblbm_b200.so
BIN       := $(PKG)/d2q9-bgk
TOOLS     := $(PKG)/gen_channel

all: $(LIB) $(BIN) $(TOOLS)

$(LIB): $(wildcard $(PKG)/csrc/*.cu $(PKG)/csrc/*.cuh) include/lbm_b200.h
	@mkdir -p build
	$(NVCC) $(NVFLAGS) -shared $(PKG)/csrc/lbm_b200.cu -o $@ -lcudart 2> build/ptxas.log || (cat build/ptxas.log; exit 1)

$(BIN): $(PKG)/host/d2q9-bgk.c include/lbm_b200.h $(LIB)
	$(HOSTCC) -std=c99 -O2 -Wall -Wextra -D_POSIX_C_SOURCE=200809L -Iinclude $(PKG)/host/d2q9-bgk.c -o $@ \
	    -L$(PKG) -llbm_b200 -Wl,-rpath,'$$ORIGIN' -lm -lpthread

$(PKG)/gen_channel: $(PKG)/host/gen_channel.c
	$(HOSTCC) -std=c99 -O2 -Wall -Wextra $< -o $@

oracle:
	$(MAKE) -C oracle all ref

GRIDS ?= 128x128 128x256 256x256 1024x1024

check: all
	$(MAKE) check-all GRIDS=128x128

check-all: all
	@mkdir -p build/check/ref
	python tests/ref_check/write_goldens.py build/check/ref $(GRIDS)
	@set -e; for g in $(GRIDS); do \
	  mkdir -p build/check/$$g; echo "== $$g =="; \
	  (cd build/check/$$g && ../../../$(BIN) ../../../tests/golden/inputs/input_$$g.params ../../../tests/golden/inputs/obstacles_$$g.dat | grep -E 'Reynolds|Compute|B200'); \
	  python tests/ref_check/check.py --ref-av-vels-file=build/check/ref/$$g.av_vels.dat \
	    --ref-final-state-file=build/check/ref/$$g.final_state.dat \
	    --av-vels-file=build/check/$$g/av_vels.dat --final-state-file=build/check/$$g/final_state.dat; \
	done

clean:
	rm -rf build $(LIB) $(BIN) $(TOOLS)

.PHONY: all oracle check check-all clean
