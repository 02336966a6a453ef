# Top-level build.  Products:
#   crackling_b200/lib/libissl_cuda.so   C-ABI library (include/issl_cuda.h), sm_100a only
#   bin/isslScoreOfftargets              drop-in host program (same CLI as the reference binary)
#   bin/isslScoreServer                  resident scorer behind the same CLI (index stays in HBM between pages)
#   bin/extractOfftargets                drop-in off-target site extractor (FASTA -> sorted site list, optionally -> .issl)
#   bin/isslCreateIndex                  drop-in index builder (same CLI, byte-identical .issl)
# Test infrastructure (never linked into the products):
#   make oracle   -> oracle/_build/libissl_oracle.so
#   make ref      -> oracle/_ref/* (the unmodified reference, compiled from /root/reference)
NVCC      ?= /usr/local/cuda/bin/nvcc
CXX       := g++
ARCH      := -gencode arch=compute_100a,code=sm_100a
CSRC      := crackling_b200/csrc
LIBDIR    := crackling_b200/lib
NVFLAGS   := $(ARCH) -O3 -std=c++17 -lineinfo -fmad=false -Xcompiler -fPIC,-fopenmp,-Wall -Iinclude -I$(CSRC) --expt-relaxed-constexpr
CXXFLAGS  := -O3 -std=c++17 -fPIC -fopenmp -Wall -Wextra -Iinclude -I$(CSRC)

all: $(LIBDIR)/libissl_cuda.so bin/isslScoreOfftargets bin/isslCreateIndex bin/isslScoreServer bin/extractOfftargets

$(LIBDIR)/issl_host.o: $(CSRC)/issl_host.cpp $(CSRC)/issl_internal.h $(CSRC)/issl_triple_tables.h include/issl_cuda.h
	@mkdir -p $(LIBDIR)
	$(CXX) $(CXXFLAGS) -c -o $@ $<

$(LIBDIR)/issl_device.o: $(CSRC)/issl_device.cu $(CSRC)/issl_device_common.cuh $(CSRC)/issl_kernels.cuh $(CSRC)/issl_triple.cuh $(CSRC)/issl_triple_tables.h $(CSRC)/issl_internal.h $(CSRC)/cfd_tables.h include/issl_cuda.h
	@mkdir -p $(LIBDIR)
	$(NVCC) $(NVFLAGS) -Xptxas -v -c -o $@ $< 2> $(LIBDIR)/ptxas.log || (cat $(LIBDIR)/ptxas.log; false)

$(LIBDIR)/issl_sites.o: $(CSRC)/issl_sites.cu $(CSRC)/issl_device_common.cuh $(CSRC)/issl_internal.h include/issl_cuda.h
	@mkdir -p $(LIBDIR)
	$(NVCC) $(NVFLAGS) -Xptxas -v -c -o $@ $< 2> $(LIBDIR)/ptxas_sites.log || (cat $(LIBDIR)/ptxas_sites.log; false)

$(LIBDIR)/issl_multi.o: $(CSRC)/issl_multi.cpp $(CSRC)/issl_internal.h include/issl_cuda.h
	@mkdir -p $(LIBDIR)
	$(CXX) $(CXXFLAGS) -c -o $@ $<

$(LIBDIR)/libissl_cuda.so: $(LIBDIR)/issl_host.o $(LIBDIR)/issl_multi.o $(LIBDIR)/issl_device.o $(LIBDIR)/issl_sites.o
	$(NVCC) $(ARCH) -shared -o $@ $^ -Xcompiler -fopenmp -lgomp -lpthread

HOSTHDR   := include/issl_cuda.h $(CSRC)/issl_hostcommon.h $(CSRC)/issl_wire.h

bin/isslScoreOfftargets: $(CSRC)/isslScoreOfftargets.cpp $(HOSTHDR) $(LIBDIR)/libissl_cuda.so
	@mkdir -p bin
	$(CXX) -O2 -std=c++17 -fopenmp -Wall -Wextra -Iinclude -I$(CSRC) -o $@ $< -L$(LIBDIR) -lissl_cuda -pthread '-Wl,-rpath,$$ORIGIN/../$(LIBDIR)'

bin/isslScoreServer: $(CSRC)/isslScoreServer.cpp $(HOSTHDR) $(LIBDIR)/libissl_cuda.so
	@mkdir -p bin
	$(CXX) -O2 -std=c++17 -Wall -Wextra -Iinclude -I$(CSRC) -o $@ $< -L$(LIBDIR) -lissl_cuda -pthread '-Wl,-rpath,$$ORIGIN/../$(LIBDIR)'

bin/isslCreateIndex: $(CSRC)/isslCreateIndex.cpp include/issl_cuda.h $(LIBDIR)/libissl_cuda.so
	@mkdir -p bin
	$(CXX) -O2 -std=c++17 -Wall -Wextra -Iinclude -o $@ $< -L$(LIBDIR) -lissl_cuda '-Wl,-rpath,$$ORIGIN/../$(LIBDIR)'

bin/extractOfftargets: $(CSRC)/extractOfftargets.cpp include/issl_cuda.h $(LIBDIR)/libissl_cuda.so
	@mkdir -p bin
	$(CXX) -O2 -std=c++17 -Wall -Wextra -Iinclude -o $@ $< -L$(LIBDIR) -lissl_cuda '-Wl,-rpath,$$ORIGIN/../$(LIBDIR)'

oracle:
	$(MAKE) -C oracle oracle
ref:
	$(MAKE) -C oracle ref

clean:
	rm -rf $(LIBDIR) bin oracle/_build

.PHONY: all oracle ref clean
