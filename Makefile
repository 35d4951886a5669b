# Convenience targets (the driver uses __graft_entry__.build() / pytest / bench.py directly).
PY ?= python

build:            ## nvcc -gencode arch=compute_100a,code=sm_100a -> roboticmanipulators_b200/librk_b200.so
	$(PY) -c "import __graft_entry__ as g; g.build()"

test-cpu:         ## oracle vs goldens, host logic, ABI, contract of the committed records (no GPU)
	$(PY) -m pytest tests -x -q -m "not gpu"

test-gpu:         ## parity through the C ABI (needs a B200)
	$(PY) -m pytest tests -x -q -m gpu

smoke:
	$(PY) -c "import __graft_entry__ as g; g.smoke()"

bench:            ## one JSON line; see DESIGN.md section 6
	$(PY) bench.py

bench-reference:  ## the CPU arm (NumPy restatement on all cores + a bounded sample of the unmodified reference)
	$(PY) bench.py --impl reference

goldens:          ## regenerate tests/golden/*.npz from the real reference (needs /root/reference; ~20 min)
	$(PY) tests/golden/make_golden.py
	$(PY) tests/golden/make_golden.py --only config1
	$(PY) tests/golden/make_golden.py --only general --general-n 4000
	$(PY) tests/golden/make_golden.py --only armslarge,trajlarge,chainrandom
	$(PY) tests/golden/make_optimize_golden.py

.PHONY: build test-cpu test-gpu smoke bench bench-reference goldens
