# Convenience targets; the library itself is built by egobox_b200/_build.py (nvcc, sm_100a only, in tree).
PY ?= python

.PHONY: lib test-cpu test-gpu example bench rust-ffi sass clean

lib:            ## egobox_b200/libegobox_gpu.so
	$(PY) -m egobox_b200._build

test-cpu: lib   ## oracle vs golden vectors, host logic, C-ABI exports (no GPU)
	$(PY) -m pytest tests -q -m "not gpu"

test-gpu: lib   ## CUDA path vs oracle through the C ABI (needs a B200)
	$(PY) -m pytest tests -q -m gpu

example: lib    ## the reference's crates/gp/examples/kriging.rs in C99 over the C ABI
	gcc -std=c99 -Wall -Wextra -pedantic -Iinclude examples/kriging.c -Legobox_b200 -legobox_gpu \
	    -Wl,-rpath,$(CURDIR)/egobox_b200 -lm -o build/kriging_c

bench: lib      ## one JSON line (needs a B200)
	$(PY) bench.py

rust-ffi:       ## regenerate bindings/rust/cuda_ffi.rs from include/egobox_gpu.h
	$(PY) tools/gen_rust_ffi.py

sass: lib       ## per-kernel tcgen05 / TMA / DMMA instruction counts
	$(PY) tools/sass_evidence.py

clean:
	rm -rf build egobox_b200/libegobox_gpu.so
