# Convenience targets.  The driver-facing entry points are __graft_entry__.py (build / smoke) and bench.py.
PY ?= python

.PHONY: lib oracle test-cpu test-gpu bench smoke clean
lib:            ## nvcc -gencode arch=compute_100a,code=sm_100a ... -> cuda-lbfgs_b200/lib/liblbfgsb200.so
	$(PY) cuda-lbfgs_b200/build.py
oracle:         ## test infrastructure: C restatement + the unmodified reference (needs /root/reference for the latter)
	$(MAKE) -C oracle all
test-cpu: lib oracle
	$(PY) -m pytest tests -q -m "not gpu"
test-gpu: lib oracle
	$(PY) -m pytest tests -q -m gpu
smoke: lib
	$(PY) __graft_entry__.py smoke
bench: lib
	$(PY) bench.py
clean:
	rm -rf cuda-lbfgs_b200/lib oracle/_ref oracle/liblbfgs_oracle.so
