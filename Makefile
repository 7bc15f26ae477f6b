# Builds libkbotstep.so (sm_100a only) in-tree.  `python -c "import __graft_entry__ as g; g.build()"` calls this.
NVCC ?= nvcc
ARCH := -gencode arch=compute_100a,code=sm_100a
PKG := kbot-joystick_b200
SRC := $(PKG)/csrc
OUT := $(PKG)/libkbotstep.so
COMMON := -O3 -std=c++17 -lineinfo $(ARCH) -Iinclude -I$(SRC) -Xcompiler -fPIC -Xptxas -v
OBJS := $(SRC)/kbs_api.o $(SRC)/kbs_elementwise.o $(SRC)/kbs_net_simt.o $(SRC)/kbs_net_tc.o $(SRC)/kbs_ppo_update.o

all: $(OUT)

$(SRC)/kbs_elementwise.o: $(SRC)/kbs_elementwise.cu $(SRC)/kbs_common.cuh include/kbotstep.h
	$(NVCC) $(COMMON) -fmad=false -c $< -o $@

$(SRC)/%.o: $(SRC)/%.cu $(SRC)/kbs_common.cuh include/kbotstep.h
	$(NVCC) $(COMMON) -c $< -o $@

$(OUT): $(OBJS)
	$(NVCC) $(ARCH) -shared -o $@ $(OBJS) -lcudart

clean:
	rm -f $(OBJS) $(OUT)
.PHONY: all clean

# Optional: XLA-FFI (jax.ffi) handlers over the C-ABI.  Needs jaxlib's headers, which this image does not have:
#   make ffi JAX_INCLUDE=$(python -c "import jax.ffi; print(jax.ffi.include_dir())")
ffi: $(OUT)
	@test -n "$(JAX_INCLUDE)" || (echo "set JAX_INCLUDE to jax.ffi.include_dir()"; exit 1)
	$(NVCC) -O2 -std=c++17 -x cu $(ARCH) -Iinclude -I$(JAX_INCLUDE) -Xcompiler -fPIC -shared \
	  $(SRC)/kbs_xla_ffi.cc -o $(PKG)/libkbs_xla_ffi.so -L$(PKG) -lkbotstep -Xlinker -rpath -Xlinker '$$ORIGIN' -lcudart
.PHONY: ffi

# Type-check the XLA-FFI shim without jaxlib: g++ against the stub FFI surface in tests/ffi_stub (every handler is
# static_asserted against its binding); run by tests/test_host_cpu.py.
ffi-check:
	g++ -std=c++17 -fsyntax-only -Wall -Itests/ffi_stub -Iinclude -I/usr/local/cuda/include -x c++ $(SRC)/kbs_xla_ffi.cc
.PHONY: ffi-check
