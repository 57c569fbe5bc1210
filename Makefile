# Builds the product library in-tree: locarna_b200/liblocarna_b200.so (sm_100a only).
NVCC      ?= nvcc
ARCH      := -gencode arch=compute_100a,code=sm_100a
NVCCFLAGS := $(ARCH) -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xcompiler -Wall -Xcompiler -Wno-unused-function -ccbin /usr/bin/g++
SRC       := locarna_b200/csrc
OBJ       := build/obj
LIB       := locarna_b200/liblocarna_b200.so

OBJS := $(OBJ)/kernels.o $(OBJ)/dfill_rows.o $(OBJ)/builder.o $(OBJ)/envelope.o $(OBJ)/runtime.o $(OBJ)/host_model.o $(OBJ)/guide_tree.o $(OBJ)/allpairs.o

CLI       := locarna_b200/bin/locarna_b200
CLI_P     := locarna_b200/bin/locarna_p_b200
CLI_TREE  := locarna_b200/bin/mlocarna_tree_b200

REFSRC    := /root/reference/src/locarna.cc
CLI_REF   := locarna_b200/bin/locarna_refmain_b200

all: $(LIB) $(CLI) $(CLI_P) $(CLI_TREE) $(if $(wildcard $(REFSRC)),$(CLI_REF))

$(OBJ):
	mkdir -p $(OBJ)

$(OBJ)/%.o: $(SRC)/%.cu $(wildcard $(SRC)/*.h) $(wildcard $(SRC)/*.cuh) include/locarna_b200.h | $(OBJ)
	$(NVCC) $(NVCCFLAGS) -c $< -o $@

# host-only code: plain x86-64 code generation (no -march, no fast-math) so that the 80-bit envelope
# arithmetic is evaluated exactly like the reference's
$(OBJ)/%.o: $(SRC)/%.cc $(wildcard $(SRC)/*.h) | $(OBJ)
	/usr/bin/g++ -std=c++17 -O3 -fPIC -Wall -c $< -o $@

$(LIB): $(OBJS)
	$(NVCC) $(ARCH) -shared -o $@ $(OBJS) -cudart static -lpthread

# `locarna`-compatible command line front end (host C++ over the C ABI)
$(CLI): $(SRC)/cli/locarna_main.cc include/locarna_b200.hh include/locarna_b200.h $(LIB)
	mkdir -p locarna_b200/bin
	/usr/bin/g++ -std=c++17 -O2 -Wall -Iinclude $< -o $@ -Llocarna_b200 -llocarna_b200 -Wl,-rpath,'$$ORIGIN/..'

# `locarna_p`-compatible front end (LocARNA-P: partition function, arc-match / base-match probabilities)
$(CLI_P): $(SRC)/cli/locarna_p_main.cc include/locarna_b200.hh include/locarna_b200.h $(LIB)
	mkdir -p locarna_b200/bin
	/usr/bin/g++ -std=c++17 -O2 -Wall -Iinclude $< -o $@ -Llocarna_b200 -llocarna_b200 -Wl,-rpath,'$$ORIGIN/..'

# all-vs-all guide-tree stage of mlocarna in one process (score matrix, score list, UPGMA tree), plain C ABI
$(CLI_TREE): $(SRC)/cli/mlocarna_tree_main.cc include/locarna_b200.h $(LIB)
	mkdir -p locarna_b200/bin
	/usr/bin/g++ -std=c++17 -O2 -Wall -Iinclude $< -o $@ -Llocarna_b200 -llocarna_b200 -Wl,-rpath,'$$ORIGIN/..'

# The reference's own pipeline code on the B200 library: the body of run_and_report() is cut from the reference tree (never committed)
# and compiled unmodified against include/locarna_b200_compat.hh. Built only where the reference tree is present.
refmain: $(CLI_REF)
build/refmain/refmain_block.inc: $(REFSRC) tools/extract_ref_block.py
	mkdir -p build/refmain
	python tools/extract_ref_block.py $(REFSRC) $@
$(CLI_REF): $(SRC)/cli/refmain_driver.cc build/refmain/refmain_block.inc include/locarna_b200_compat.hh include/locarna_b200.hh include/locarna_b200.h $(LIB)
	mkdir -p locarna_b200/bin
	/usr/bin/g++ -std=c++17 -O2 -Wall -Wno-unused-variable -Wno-unused-local-typedefs -Iinclude -Ibuild/refmain $< -o $@ -Llocarna_b200 -llocarna_b200 -Wl,-rpath,'$$ORIGIN/..'

clean:
	rm -rf build $(LIB) locarna_b200/bin
