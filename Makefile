# Top-level build. `make` builds everything; `make host` only the CPU-side pieces
# (synthetic generator, oracle port) that the non-GPU tests need.
NVCC      ?= /usr/local/cuda/bin/nvcc
CC        ?= gcc
ARCH      := -gencode arch=compute_100a,code=sm_100a
NVFLAGS   := -O3 -std=c++17 -lineinfo --extended-lambda $(ARCH) -Xcompiler -fPIC -Xcompiler -Wall -Iinclude -Icbc_b200/csrc
BUILD     := cbc_b200/_build
CU_SRC    := $(wildcard cbc_b200/csrc/*.cu)
CU_HDR    := $(wildcard cbc_b200/csrc/*.cuh) $(wildcard cbc_b200/csrc/*.h) $(wildcard include/*.h)
HOST_SRC  := cbc_b200/csrc/host/sam_ingest.c
HOST_HDR  := $(wildcard cbc_b200/csrc/host/*.h) $(wildcard include/*.h)

.PHONY: all host cuda cli oracle clean
all: host cuda cli

host: $(BUILD)/libcbcsynth.so $(BUILD)/libcbchost.so oracle

oracle:
	@$(MAKE) -s -C oracle port
	@if [ -d /root/reference/src ] && [ ! -x oracle/_ref/cbc_ref ]; then $(MAKE) -s -C oracle ref; fi

$(BUILD)/libcbcsynth.so: cbc_b200/csrc/host/synth.c cbc_b200/csrc/host/synth.h
	@mkdir -p $(BUILD)
	$(CC) -O2 -Wall -fPIC -shared $< -o $@

$(BUILD)/libcbchost.so: $(HOST_SRC) $(HOST_HDR)
	@mkdir -p $(BUILD)
	$(CC) -O2 -Wall -fPIC -shared -Iinclude -Icbc_b200/csrc/host $(HOST_SRC) -o $@ -lpthread

cuda: $(BUILD)/libcbcg.so

$(BUILD)/libcbcg.so: $(CU_SRC) $(CU_HDR)
	@mkdir -p $(BUILD)
	$(NVCC) $(NVFLAGS) -shared $(CU_SRC) -o $@ -lcudart

cli: $(BUILD)/cbc

$(BUILD)/cbc: cbc_b200/csrc/host/cbc_main.c $(HOST_SRC) $(HOST_HDR) $(BUILD)/libcbcg.so
	$(CC) -O2 -Wall -Iinclude -Icbc_b200/csrc/host cbc_b200/csrc/host/cbc_main.c $(HOST_SRC) \
	    -L$(BUILD) -lcbcg -Wl,-rpath,'$$ORIGIN' -lpthread -o $@

clean:
	rm -rf $(BUILD) oracle/_build
