# Top-level build: everything in-tree, sm_100a only.
#   make            -> host library, CUDA library, C++ drop-in app, oracle port (+ oracle/_ref when
#                      /root/reference is present)
#   make cuda|host|app|oracle
NVCC     := /usr/local/cuda/bin/nvcc
CXX      := /usr/bin/g++
ARCH     := -gencode arch=compute_100a,code=sm_100a
BUILD    := raytracert_b200/_build
HOSTDIR  := raytracert_b200/host
CSRC     := raytracert_b200/csrc
HOSTFLAGS := -std=c++17 -O2 -ffp-contract=off -fPIC -Wall -Wextra -Iinclude
# no -use_fast_math / -ftz: the exact path needs IEEE binary32 with denormals ("parity")
NVFLAGS  := $(ARCH) -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Iinclude --fmad=true

.PHONY: all host cuda app oracle clean
all: host cuda app oracle

host: $(BUILD)/librt_host.so
cuda: $(BUILD)/librt_b200.so
app:  $(BUILD)/rt_main
oracle:
	$(MAKE) -C oracle all

$(BUILD)/librt_host.so: $(HOSTDIR)/mesh.cpp $(HOSTDIR)/host_capi.cpp $(wildcard $(HOSTDIR)/*.h) include/rt_b200.h
	@mkdir -p $(BUILD)
	$(CXX) $(HOSTFLAGS) -shared -o $@ $(HOSTDIR)/mesh.cpp $(HOSTDIR)/host_capi.cpp

$(BUILD)/librt_b200.so: $(wildcard $(CSRC)/*.cu) $(wildcard $(CSRC)/*.cuh) $(wildcard $(CSRC)/*.h) include/rt_b200.h
	@mkdir -p $(BUILD)
	$(NVCC) $(NVFLAGS) -shared -o $@ $(CSRC)/rt_b200.cu -cudart static -ldl

$(BUILD)/rt_main: $(HOSTDIR)/main.cpp $(HOSTDIR)/raytracing.cpp $(HOSTDIR)/mesh.cpp $(wildcard $(HOSTDIR)/*.h) include/rt_b200.h $(BUILD)/librt_b200.so
	$(CXX) $(HOSTFLAGS) -o $@ $(HOSTDIR)/main.cpp $(HOSTDIR)/raytracing.cpp $(HOSTDIR)/mesh.cpp -L$(BUILD) -lrt_b200 -Wl,-rpath,'$$ORIGIN'

clean:
	rm -rf $(BUILD)
	$(MAKE) -C oracle clean
