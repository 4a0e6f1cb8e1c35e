# Builds the product library (sm_100a only), the tuner, and the oracle pieces.
#   make            -> procedural-universe_b200/lib/libnbody_b200.so, procedural-universe_b200/bin/nbody_headless
#   make tools      -> tools/tune_allpairs
#   make oracle     -> oracle/_build/libnbody_port.so (+ oracle/_ref/libpu_ref.so when /root/reference exists)
NVCC ?= /usr/local/cuda/bin/nvcc
CXX ?= g++
ARCH := -gencode arch=compute_100a,code=sm_100a
PKG := procedural-universe_b200
SRC := $(PKG)/csrc
OBJ := $(PKG)/build
LIB := $(PKG)/lib/libnbody_b200.so

NVCCFLAGS := $(ARCH) -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xcompiler -fvisibility=hidden
# -fmad stays on for device code (the kernels state where fusion matters); host seeder must not fuse.
CXXFLAGS := -O2 -std=c++17 -fPIC -ffp-contract=off -fvisibility=hidden -I/usr/local/cuda/include

CU_SRCS := $(SRC)/nb_api.cu $(SRC)/integrate.cu $(SRC)/tree.cu $(SRC)/energy.cu $(SRC)/probe.cu $(SRC)/seed_device.cu $(SRC)/p2p.cu $(SRC)/query.cu
CPP_SRCS := $(SRC)/seed_host.cpp $(SRC)/nccl_dl.cpp $(SRC)/nbody_io.cpp
CU_OBJS := $(patsubst $(SRC)/%.cu,$(OBJ)/%.o,$(CU_SRCS))
CPP_OBJS := $(patsubst $(SRC)/%.cpp,$(OBJ)/%.o,$(CPP_SRCS))
HDRS := $(wildcard $(SRC)/*.h $(SRC)/*.cuh) include/nbody_b200.h

BIN := $(PKG)/bin/nbody_headless

all: $(LIB) $(BIN)

$(BIN): $(PKG)/host/nbody_headless.cpp include/nbody_b200.h $(LIB)
	@mkdir -p $(PKG)/bin
	$(CXX) -O2 -std=c++17 -Iinclude -o $@ $< -L$(PKG)/lib -lnbody_b200 -Wl,-rpath,'$$ORIGIN/../lib'

$(OBJ)/%.o: $(SRC)/%.cu $(HDRS)
	@mkdir -p $(OBJ)
	$(NVCC) $(NVCCFLAGS) -c $< -o $@

$(OBJ)/%.o: $(SRC)/%.cpp $(HDRS)
	@mkdir -p $(OBJ)
	$(CXX) $(CXXFLAGS) -c $< -o $@

$(LIB): $(CU_OBJS) $(CPP_OBJS)
	@mkdir -p $(PKG)/lib
	$(NVCC) $(ARCH) -shared -o $@ $^ -cudart static -ldl -lpthread

tools: tools/tune_allpairs tools/issue_probe

tools/issue_probe: tools/issue_probe.cu
	$(NVCC) $(ARCH) -O3 -o $@ $<

tools/tune_allpairs: tools/tune_allpairs.cu $(SRC)/allpairs.cuh
	$(NVCC) $(ARCH) -O3 -lineinfo -std=c++17 -o $@ $<

oracle:
	$(MAKE) -C oracle

clean:
	rm -rf $(OBJ) $(PKG)/lib $(PKG)/bin tools/tune_allpairs tools/issue_probe

.PHONY: all tools oracle clean
