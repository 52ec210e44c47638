# Top-level build: libndb_b200.so (sm_100a only) + the oracle (test infrastructure).
NVCC      ?= /usr/local/cuda/bin/nvcc
HOSTCXX   ?= /usr/bin/g++
ARCH      := -gencode arch=compute_100a,code=sm_100a
NVFLAGS   := $(ARCH) -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xcompiler -Wall -ccbin $(HOSTCXX) \
             --expt-relaxed-constexpr -Xptxas -warn-spills
SRCDIR    := neurondb_b200/csrc
LIBDIR    := neurondb_b200/lib
OBJDIR    := build/obj
SRCS      := $(wildcard $(SRCDIR)/*.cu)
OBJS      := $(patsubst $(SRCDIR)/%.cu,$(OBJDIR)/%.o,$(SRCS))
HDRS      := $(wildcard $(SRCDIR)/*.cuh) include/ndb_b200.h

all: $(LIBDIR)/libndb_b200.so oracle glue

$(OBJDIR)/%.o: $(SRCDIR)/%.cu $(HDRS)
	@mkdir -p $(OBJDIR)
	$(NVCC) $(NVFLAGS) -c $< -o $@

$(LIBDIR)/libndb_b200.so: $(OBJS)
	@mkdir -p $(LIBDIR)
	$(NVCC) $(ARCH) -shared -ccbin $(HOSTCXX) -o $@ $(OBJS) -ldl

oracle:
	$(MAKE) -s -C oracle

# reference-side glue (integration/*.c) against the reference's own header; needs the library first
glue: $(LIBDIR)/libndb_b200.so
	$(MAKE) -s -C oracle glue

# development build with epilogue statistics in tc_knn_kernel (tools/c4_probe.py; NDB_B200_LIB_PATH selects it)
counters: $(LIBDIR)/libndb_b200.so
	@mkdir -p $(OBJDIR)/ctr
	$(NVCC) $(NVFLAGS) -DNDB_TC_COUNTERS -c $(SRCDIR)/tc_knn.cu -o $(OBJDIR)/ctr/tc_knn.o
	$(NVCC) $(ARCH) -shared -ccbin $(HOSTCXX) -o $(LIBDIR)/libndb_b200_ctr.so $(filter-out $(OBJDIR)/tc_knn.o,$(OBJS)) $(OBJDIR)/ctr/tc_knn.o -ldl

clean:
	rm -rf build $(LIBDIR)/*.so
	$(MAKE) -C oracle clean

.PHONY: all oracle glue clean counters
