# Top-level build: product library (CUDA, sm_100a), synthetic workload generator, CPU checkers.
NVCC ?= /usr/local/cuda/bin/nvcc
PKG = gmap-gsnap_b200
NVFLAGS = -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC,-O3,-Wall,-Wno-unused-function -Xptxas -v

all: cuda synth oracle

cuda: $(PKG)/csrc/libdynprog_cuda.so
$(PKG)/csrc/libdynprog_cuda.so: $(PKG)/csrc/*.cu $(PKG)/csrc/*.h include/dynprog_cuda.h
	$(NVCC) $(NVFLAGS) -shared -o $@ $(PKG)/csrc/dynprog_cuda.cu -lcudart 2> $(PKG)/csrc/ptxas.log || (cat $(PKG)/csrc/ptxas.log; exit 1)
	@grep -E "registers|spill" $(PKG)/csrc/ptxas.log | sort | uniq -c | head -40

synth: $(PKG)/host/libdpc_synth.so
$(PKG)/host/libdpc_synth.so: $(PKG)/host/synth.c include/dynprog_cuda.h
	gcc -O2 -fPIC -Wall -shared -o $@ $(PKG)/host/synth.c -lm

oracle:
	$(MAKE) -C oracle

clean:
	rm -f $(PKG)/csrc/*.so $(PKG)/host/*.so $(PKG)/csrc/ptxas.log
	$(MAKE) -C oracle clean
.PHONY: all cuda synth oracle clean
