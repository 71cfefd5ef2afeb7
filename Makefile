# Builds the sm_100a C-ABI library, the CPU oracle (test infrastructure) and, when /root/reference is mounted,
# oracle/_ref. `python -c "import __graft_entry__ as g; g.build()"` drives this.
NVCC      ?= nvcc
CXX       ?= g++
ARCH      := -gencode arch=compute_100a,code=sm_100a
NVCCFLAGS := $(ARCH) -O3 -std=c++17 -lineinfo -Xcompiler -fPIC,-Wall,-Wno-unused-function -Xptxas -v
CSRC      := hyrise_b200/csrc
LIB       := hyrise_b200/lib/libhyrise_b200.so
CU_SRCS   := $(wildcard $(CSRC)/*.cu)
CU_OBJS   := $(patsubst $(CSRC)/%.cu,build/%.o,$(CU_SRCS))
HDRS      := $(wildcard $(CSRC)/*.hpp) $(wildcard $(CSRC)/*.cuh) include/hyrise_b200.h

TPCH_LIB  := hyrise_b200/lib/libhyb_tpch.so

all: $(LIB) $(TPCH_LIB) oracle

build/%.o: $(CSRC)/%.cu $(HDRS)
	@mkdir -p build
	$(NVCC) $(NVCCFLAGS) -c $< -o $@ 2> build/$*.ptxas.log || (cat build/$*.ptxas.log; exit 1)

$(LIB): $(CU_OBJS)
	@mkdir -p hyrise_b200/lib
	$(NVCC) $(ARCH) -shared -o $@ $(CU_OBJS) -cudart shared

$(TPCH_LIB): $(CSRC)/tpch_gen.cpp include/hyrise_b200_tpch.h include/hyrise_b200.h
	@mkdir -p hyrise_b200/lib
	$(CXX) -O3 -march=x86-64-v3 -std=c++17 -fPIC -Wall -pthread -shared -o $@ $(CSRC)/tpch_gen.cpp

oracle:
	$(MAKE) -C oracle

# The C++ host-side mirror of the reference operator interface (include/hyrise_b200.hpp) driving the three operators.
example: build/tpch_operators
build/tpch_operators: examples/tpch_operators.cpp include/hyrise_b200.hpp include/hyrise_b200.h include/hyrise_b200_tpch.h $(LIB) $(TPCH_LIB)
	@mkdir -p build
	$(CXX) -O2 -std=c++17 -Wall -Wextra -Iinclude -o $@ examples/tpch_operators.cpp -Lhyrise_b200/lib -lhyrise_b200 -lhyb_tpch \
		-Wl,-rpath,'$$ORIGIN/../hyrise_b200/lib'

# Host-only check of the mirror's predicate normalisation against the reference's own test expectations.
build/predicate_cast_check: tests/cpp/predicate_cast_check.cpp include/hyrise_b200.hpp include/hyrise_b200.h $(LIB)
	@mkdir -p build
	$(CXX) -O1 -std=c++17 -Wall -Wextra -Iinclude -o $@ tests/cpp/predicate_cast_check.cpp -Lhyrise_b200/lib -lhyrise_b200 \
		-Wl,-rpath,'$$ORIGIN/../hyrise_b200/lib'

clean:
	rm -rf build $(LIB) $(TPCH_LIB)
	$(MAKE) -C oracle clean

.PHONY: all oracle clean example
