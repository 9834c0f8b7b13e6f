# Top-level convenience targets with the reference's names (reference Makefile:61-68): `make test` builds and runs
# the -t harness, `make memcheck` runs it under compute-sanitizer.  When the reference tree is present (REF) the
# harness is the reference's OWN src/test.cu + src/main.cpp, compiled unmodified against include/compat and linked
# with librmd_compat.a (oracle/Makefile: _ref/ref_harness); otherwise it is examples/harness.cpp, the same flow on
# the C ABI.  Everything is built for sm_100a only.
PYTHON ?= python
NVCC   ?= /usr/local/cuda/bin/nvcc
REF    ?= /root/reference
BUILD  := build
LIBDIR := raymarchdenoisercuda_b200

.PHONY: all lib test memcheck clean

all: lib

lib:
	$(PYTHON) -c "import __graft_entry__ as g; g.build()"

$(BUILD)/main: lib
	mkdir -p $(BUILD)
	@if [ -d $(REF)/src ]; then \
	    $(MAKE) -s -C oracle _ref/ref_harness REF=$(REF) && cp oracle/_ref/ref_harness $@; \
	 elif [ -x oracle/_ref/ref_harness ]; then cp oracle/_ref/ref_harness $@; \
	 else g++ -std=c++17 -Iinclude -I/usr/local/cuda/include examples/harness.cpp -o $@ -L$(LIBDIR) -lrmd_b200 \
	      -L/usr/local/cuda/lib64 -lcudart -Wl,-rpath,$(CURDIR)/$(LIBDIR) -Wl,-rpath,/usr/local/cuda/lib64; fi

test: $(BUILD)/main
	@./$(BUILD)/main -t

memcheck: $(BUILD)/main
	compute-sanitizer --tool memcheck --show-backtrace=yes --log-file $(BUILD)/memcheck.log ./$(BUILD)/main -t

clean:
	rm -rf $(BUILD) examples/build $(LIBDIR)/csrc/build $(LIBDIR)/*.so $(LIBDIR)/*.a
	$(MAKE) -s -C oracle clean
