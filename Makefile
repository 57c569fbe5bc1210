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

all: $(LIB) $(CLI) $(CLI_P) $(CLI_TREE)

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

clean:
	rm -rf build $(LIB) locarna_b200/bin
