# Builds libvfa_b200.so (sm_100a only) in-tree.  `python -c "import __graft_entry__ as g; g.build()"` calls this.
NVCC      ?= nvcc
ARCH      := -gencode arch=compute_100a,code=sm_100a
NVCCFLAGS := -O3 -std=c++17 -lineinfo $(ARCH) -Xcompiler -fPIC,-Wall -Xptxas -v
SRC_DIR   := vfa_b200/csrc
OUT       := vfa_b200/lib/libvfa_b200.so
SRCS      := $(wildcard $(SRC_DIR)/*.cu)
OBJS      := $(patsubst $(SRC_DIR)/%.cu,build/%.o,$(SRCS))

all: $(OUT)

build/%.o: $(SRC_DIR)/%.cu $(wildcard $(SRC_DIR)/*.cuh) include/vfa_b200.h
	@mkdir -p build
	$(NVCC) $(NVCCFLAGS) $(if $(filter vfa_table,$*),-fmad=false,) -c $< -o $@ 2> build/$*.ptxas.log || (cat build/$*.ptxas.log; false)

$(OUT): $(OBJS)
	@mkdir -p vfa_b200/lib
	$(NVCC) $(ARCH) -shared -o $@ $^ -cudart static -ldl

clean:
	rm -rf build $(OUT)
.PHONY: all clean
