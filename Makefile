# Builds liboptados_b200.so (CUDA, sm_100a only) and the CPU oracle (test infrastructure).
NVCC      ?= nvcc
ARCH      := -gencode arch=compute_100a,code=sm_100a
NVCCFLAGS := -O3 -std=c++17 -lineinfo $(ARCH) -Xcompiler -fPIC,-Wall,-Wno-unused-function -Xptxas -v --expt-relaxed-constexpr
CSRC      := optados_b200/csrc
SRCS      := $(CSRC)/od_ctx.cu $(CSRC)/od_spectral.cu $(CSRC)/od_weights.cu $(CSRC)/od_drivers.cu $(CSRC)/od_util.cu $(CSRC)/od_narrow.cu $(CSRC)/od_post.cu $(CSRC)/od_fixedwide.cu $(CSRC)/od_tasks.cu $(CSRC)/od_analysis.cu
HDRS      := $(wildcard $(CSRC)/*.h $(CSRC)/*.cuh) include/optados_b200.h
OBJS      := $(SRCS:.cu=.o) $(CSRC)/od_io.o $(CSRC)/od_run.o
LIB       := optados_b200/lib/liboptados_b200.so

all: $(LIB) oracle

$(CSRC)/%.o: $(CSRC)/%.cu $(HDRS)
	$(NVCC) $(NVCCFLAGS) -c $< -o $@ 2> $@.ptxas.log || (cat $@.ptxas.log; exit 1)

$(CSRC)/od_io.o: $(CSRC)/od_io.cpp include/optados_b200.h
	$(CXX) -O2 -std=c++17 -fPIC -Wall -c $< -o $@

$(CSRC)/od_run.o: $(CSRC)/od_run.cpp include/optados_b200.h
	$(CXX) -O2 -std=c++17 -fPIC -Wall -c $< -o $@

$(LIB): $(OBJS)
	@mkdir -p optados_b200/lib
	$(NVCC) $(ARCH) -shared -o $@ $(OBJS) -ldl

oracle:
	$(MAKE) -s -C oracle

clean:
	rm -f $(OBJS) $(CSRC)/*.ptxas.log $(LIB)
	$(MAKE) -s -C oracle clean

.PHONY: all oracle clean
