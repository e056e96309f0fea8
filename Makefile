# Builds the product library (hand-written CUDA for sm_100a behind the C ABI of include/mage_b200.h).
NVCC ?= nvcc
ARCH := -gencode arch=compute_100a,code=sm_100a
NVCCFLAGS := $(ARCH) -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xcompiler -Wall -Xptxas -v --fmad=true
SRC := mageslam_b200/csrc
OBJS := $(SRC)/capi.o $(SRC)/orb.o $(SRC)/match.o $(SRC)/ba.o $(SRC)/frontend.o $(SRC)/radius.o $(SRC)/project.o $(SRC)/undistort.o
LIB := mageslam_b200/libmage_b200.so

all: $(LIB)

$(SRC)/%.o: $(SRC)/%.cu $(SRC)/common.cuh $(SRC)/dense_ldlt.cuh include/mage_b200.h
	$(NVCC) $(NVCCFLAGS) -c $< -o $@ 2> $@.ptxas.log || (cat $@.ptxas.log; exit 1)

$(LIB): $(OBJS)
	$(NVCC) $(ARCH) -shared -o $@ $(OBJS) -lcudart_static -lpthread -ldl -lrt

oracle:
	$(MAKE) -C oracle

clean:
	rm -f $(SRC)/*.o $(SRC)/*.ptxas.log $(LIB)

.PHONY: all oracle clean
