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
