# Top-level build: the product library (tryraytrace_b200/lib/libtrt_b200.so), the test
# oracles (oracle/Makefile) and the staged assets.  Everything is compiled for sm_100a.
NVCC      ?= nvcc
CXX       ?= g++
ARCH      := -gencode arch=compute_100a,code=sm_100a
CSRC      := tryraytrace_b200/csrc
LIBDIR    := tryraytrace_b200/lib
OBJDIR    := build/obj
INC       := -Iinclude -I$(CSRC) -I$(CSRC)/kernels -I/usr/local/cuda/include
NVCCFLAGS := -O3 $(ARCH) --use_fast_math -lineinfo -std=c++17 $(INC) -Xcompiler -fPIC,-fopenmp $(NVCCFLAGS_EXTRA)
CXXFLAGS  := -O3 -march=x86-64-v3 -fopenmp -fPIC -std=c++17 $(INC) -Wall -Wno-unknown-pragmas

HOST_SRC  := loader bvh scene camera image_io pipeline renderer xorwow_tables wide_bvh
HOST_OBJS := $(addprefix $(OBJDIR)/,$(addsuffix .o,$(HOST_SRC)))
CU_OBJS   := $(OBJDIR)/wavefront.o $(OBJDIR)/bvh_build.o $(OBJDIR)/trt_capi.o

.PHONY: all lib oracle assets clean peaks
all: lib oracle assets peaks

# machine-peak microbenchmarks for the roofline (FP32 FMA rate, L2-resident read bandwidth); bench.py runs it
peaks: build/peaks
build/peaks: tools/peaks.cu
	@mkdir -p build
	$(NVCC) -O3 $(ARCH) -o $@ $<

lib: $(LIBDIR)/libtrt_b200.so $(LIBDIR)/libtrt_b200_mgpu.so

$(OBJDIR)/%.o: $(CSRC)/host/%.cpp $(wildcard include/*.h) $(wildcard $(CSRC)/host/*.h)
	@mkdir -p $(OBJDIR)
	$(CXX) $(CXXFLAGS) -c $< -o $@

$(OBJDIR)/wavefront.o: $(CSRC)/kernels/wavefront.cu $(wildcard $(CSRC)/kernels/*.cuh) include/trt_capi.h
	@mkdir -p $(OBJDIR)
	$(NVCC) $(NVCCFLAGS) -Xptxas -v -c $< -o $@ 2> $(OBJDIR)/wavefront.ptxas.log || (cat $(OBJDIR)/wavefront.ptxas.log; false)

$(OBJDIR)/bvh_build.o: $(CSRC)/kernels/bvh_build.cu $(wildcard $(CSRC)/kernels/*.cuh)
	@mkdir -p $(OBJDIR)
	$(NVCC) $(NVCCFLAGS) -c $< -o $@

$(OBJDIR)/trt_capi.o: $(CSRC)/capi/trt_capi.cu $(wildcard $(CSRC)/kernels/*.cuh) $(wildcard $(CSRC)/host/*.h) $(wildcard include/*.h)
	@mkdir -p $(OBJDIR)
	$(NVCC) $(NVCCFLAGS) -c $< -o $@

$(LIBDIR)/libtrt_b200.so: $(HOST_OBJS) $(CU_OBJS)
	@mkdir -p $(LIBDIR)
	$(NVCC) -shared $(ARCH) -o $@ $^ -Xcompiler -fopenmp -Xlinker -Bsymbolic -lgomp -ldl

# single-process multi-GPU layer (include/trt_mgpu.h): its own library, because it links NCCL
$(LIBDIR)/libtrt_b200_mgpu.so: $(CSRC)/capi/trt_mgpu.cpp include/trt_mgpu.h include/trt_capi.h $(LIBDIR)/libtrt_b200.so
	$(CXX) -O2 -fPIC -std=c++17 $(INC) -Wall -shared -o $@ $< -L$(LIBDIR) -ltrt_b200 -L/usr/local/cuda/lib64 -lcudart -lnccl -lpthread -Wl,-rpath,'$$ORIGIN'

oracle:
	$(MAKE) -C oracle all

assets:
	$(MAKE) -C oracle assets

clean:
	rm -rf build $(LIBDIR)/*.so
	$(MAKE) -C oracle clean
