# Builds libzk_b200.so (sm_100a only) in-tree, plus the C oracle helpers.
NVCC      ?= nvcc
CSRC      := zenker_audio_detection_b200/csrc
LIBDIR    := zenker_audio_detection_b200/lib
NVFLAGS   := -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -Iinclude -I$(CSRC) \
             --expt-relaxed-constexpr -Xptxas -v
SRCS      := $(CSRC)/zk_host.cu $(CSRC)/zk_gemm.cu $(CSRC)/zk_attn.cu $(CSRC)/zk_attn_split.cu $(CSRC)/zk_ops.cu $(CSRC)/zk_frontend.cu $(CSRC)/zk_model.cu $(CSRC)/zk_cascade.cu
OBJS      := $(patsubst $(CSRC)/%.cu,build/%.o,$(SRCS))
HDRS      := include/zk_b200.h $(wildcard $(CSRC)/*.cuh)

all: $(LIBDIR)/libzk_b200.so

build/%.o: $(CSRC)/%.cu $(HDRS)
	@mkdir -p build
	$(NVCC) $(NVFLAGS) -c $< -o $@ 2> build/$*.ptxas.log || (cat build/$*.ptxas.log; exit 1)

$(LIBDIR)/libzk_b200.so: $(OBJS)
	@mkdir -p $(LIBDIR)
	$(NVCC) -shared -o $@ $(OBJS) -cudart shared

# A plain C99 host that drives the path through the C ABI alone (no Python, no torch); needs a B200 to RUN.
CUDA_HOME ?= /usr/local/cuda
example: build/cascade_host
build/cascade_host: examples/cascade_host.c include/zk_b200.h $(LIBDIR)/libzk_b200.so
	@mkdir -p build
	gcc -std=c99 -O2 -Wall -Wextra -pedantic -Iinclude -isystem $(CUDA_HOME)/include $< -o $@ \
	    -L$(LIBDIR) -lzk_b200 -L$(CUDA_HOME)/lib64 -lcudart -lm \
	    -Wl,-rpath,'$$ORIGIN/../$(LIBDIR)' -Wl,-rpath,$(CUDA_HOME)/lib64

clean:
	rm -rf build $(LIBDIR)/libzk_b200.so

.PHONY: all clean example
