# Builds addapt_b200/libaddapt_b200.so (the C-ABI engine, sm_100a only), the C++ host facade + its test
# runner, the CPU oracle (test infrastructure) and the host emulation of the kernel bodies (tests only).
NVCC      ?= nvcc
ARCH      := -gencode arch=compute_100a,code=sm_100a
NVFLAGS   := $(ARCH) -lineinfo -O3 -std=c++17 -Xcompiler -fPIC,-Wall,-Wno-unused-function
CSRC      := addapt_b200/csrc
LIB       := addapt_b200/libaddapt_b200.so
PAR       := $(abspath params/rna_turner2004_reconstructed.par)
HDRS      := $(wildcard $(CSRC)/*.h) $(wildcard $(CSRC)/*.cuh) $(wildcard $(CSRC)/*.inc) include/addapt_b200.h

all: $(LIB) host oracle emu

build/%.o: $(CSRC)/%.cu $(HDRS)
	@mkdir -p build
	$(NVCC) $(NVFLAGS) -DADP_DEFAULT_PAR='"$(PAR)"' -c $< -o $@
build/energy_model.o: $(CSRC)/energy_model.cc $(HDRS)
	@mkdir -p build
	$(NVCC) $(NVFLAGS) -c $< -o $@

$(LIB): build/engine.o build/fold_kernels.o build/design_kernels.o build/energy_model.o
	$(NVCC) $(ARCH) -shared -o $@ $^ -cudart shared -Xcompiler -static-libstdc++,-static-libgcc -Xlinker --exclude-libs,ALL

host: $(LIB)
	$(MAKE) -C addapt_b200/host
oracle:
	$(MAKE) -C oracle
emu:
	$(MAKE) -C tests/emu

clean:
	rm -rf build $(LIB)
	$(MAKE) -C oracle clean
	$(MAKE) -C tests/emu clean
	$(MAKE) -C addapt_b200/host clean
.PHONY: all host oracle emu clean
