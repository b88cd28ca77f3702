# Top-level build: product library (CUDA, sm_100a), synthetic workload generator, CPU checkers.
NVCC ?= /usr/local/cuda/bin/nvcc
PKG = gmap-gsnap_b200
NVFLAGS = -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC,-O3,-Wall,-Wno-unused-function -Xptxas -v

all: cuda synth oracle emul

CUDA_SO ?= $(PKG)/csrc/libdynprog_cuda.so
cuda: $(CUDA_SO)
# DPC_DEFS: extra -D flags for tuning experiments (e.g. make cuda DPC_DEFS=-DDPC_MIN_BLOCKS_NARROW=4 CUDA_SO=/tmp/x.so)
$(CUDA_SO): $(PKG)/csrc/*.cu $(PKG)/csrc/*.h include/dynprog_cuda.h
	@mkdir -p build
	$(NVCC) $(NVFLAGS) $(DPC_DEFS) -shared -o $@ $(PKG)/csrc/dynprog_cuda.cu -lcudart 2> build/ptxas.log || (cat build/ptxas.log; exit 1)
	@grep -E "registers|spill" build/ptxas.log | sort | uniq -c | head -40

synth: $(PKG)/host/libdpc_synth.so
$(PKG)/host/libdpc_synth.so: $(PKG)/host/synth.c include/dynprog_cuda.h
	gcc -O2 -fPIC -Wall -shared -o $@ $(PKG)/host/synth.c -lm

oracle:
	$(MAKE) -C oracle

# single-lane CPU build of the device routines: test scaffolding only (tests/emul/dpc_emul.cpp)
emul: tests/emul/libdpc_emul.so tests/emul/_mock/libdynprog_cuda.so
# test double of the ticket API over the compiled reference (host-side scheduling tests without a GPU)
tests/emul/_mock/libdynprog_cuda.so: tests/emul/dpc_mock_ref.c include/dynprog_cuda.h
	mkdir -p tests/emul/_mock
	gcc -O2 -fPIC -Wall -shared -o $@ tests/emul/dpc_mock_ref.c -ldl
tests/emul/libdpc_emul.so: tests/emul/dpc_emul.cpp $(PKG)/csrc/dpc_core.h $(PKG)/csrc/dpc_host.h include/dynprog_cuda.h
	g++ -O2 -fPIC -Wall -Wextra -Wno-unknown-pragmas -shared -o $@ tests/emul/dpc_emul.cpp

# reference gmap and gmap with the drop-in solvers (BASELINE config 1); needs /root/reference
gmap: cuda
	bash oracle/build_gmap.sh

clean:
	rm -f $(PKG)/csrc/*.so $(PKG)/host/*.so build/ptxas.log tests/emul/*.so
	$(MAKE) -C oracle clean
.PHONY: all cuda synth oracle emul gmap clean
