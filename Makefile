# Builds the B200 (sm_100a) k-mer counting library, its C++ host CLI and the test oracle.
#   make            -> tsxcount_b200/lib/libtsxcuda.so + tsxcount_b200/bin/tsxcount + oracle
#   make lib | cli | oracle
NVCC     ?= /usr/local/cuda/bin/nvcc
HOSTCXX  := $(shell [ -x /usr/bin/g++ ] && echo /usr/bin/g++ || echo g++)
ARCH     := -gencode arch=compute_100a,code=sm_100a
EXTRA_DEFS ?=
NVFLAGS  := $(EXTRA_DEFS) $(ARCH) -O3 -std=c++17 -lineinfo -Xcompiler -fPIC,-Wall,-Wno-unused-function,-Wno-unknown-pragmas -ccbin $(HOSTCXX) -cudart static
CSRC     := tsxcount_b200/csrc
HOST     := tsxcount_b200/host
LIBDIR   := tsxcount_b200/lib
BINDIR   := tsxcount_b200/bin
LIB      := $(LIBDIR)/libtsxcuda.so
CLI      := $(BINDIR)/tsxcount
INGEST   := $(BINDIR)/ingest_check

.PHONY: all lib cli oracle clean
all: lib cli oracle

lib: $(LIB)

$(LIB): $(CSRC)/tsx_api.cu $(CSRC)/tsx_host_pack.cpp $(wildcard $(CSRC)/*.cuh) include/tsxcount_cuda.h
	@mkdir -p $(LIBDIR)
	$(NVCC) $(NVFLAGS) -shared -o $@ $(CSRC)/tsx_api.cu $(CSRC)/tsx_host_pack.cpp

cli: $(CLI) $(INGEST)

$(INGEST): $(HOST)/tools/ingest_check.cpp $(HOST)/FastxReader.h include/tsxcount_cuda.h $(LIB)
	@mkdir -p $(BINDIR)
	$(HOSTCXX) -O2 -std=c++17 -Wall -Iinclude -I$(HOST) -o $@ $(HOST)/tools/ingest_check.cpp -L$(LIBDIR) -ltsxcuda -lz -Wl,-rpath,'$$ORIGIN/../lib'

$(CLI): $(wildcard $(HOST)/*.cpp) $(wildcard $(HOST)/*.h) include/tsxcount_cuda.h $(LIB)
	@mkdir -p $(BINDIR)
	$(HOSTCXX) -O2 -std=c++17 -Wall -pthread -Iinclude -I/usr/local/cuda/include -o $@ $(wildcard $(HOST)/*.cpp) -L$(LIBDIR) -ltsxcuda -lnccl -lz -Wl,-rpath,'$$ORIGIN/../lib'

oracle:
	$(MAKE) -C oracle all

clean:
	rm -rf $(LIBDIR) $(BINDIR)
	$(MAKE) -C oracle clean

# A/B builds of the kernels for experiments (not shipped): make variant NAME=match EXTRA_DEFS=-DTSX_RANK_MATCH
# -> tsxcount_b200/lib/libtsxcuda_$(NAME).so, picked up with TSXC_LIB=<path> (tsxcount_b200/_lib.py)
variant:
	@mkdir -p $(LIBDIR)
	$(NVCC) $(NVFLAGS) -shared -o $(LIBDIR)/libtsxcuda_$(NAME).so $(CSRC)/tsx_api.cu $(CSRC)/tsx_host_pack.cpp
