# Convenience wrapper: the CUDA library (sm_100a), the C driver and the CPU checkers.
# `python -c "import __graft_entry__ as g; g.build()"` does the same.
all:
	$(MAKE) -C starrynight_b200/csrc
	$(MAKE) -C driver
	$(MAKE) -C oracle

test-cpu: all
	python -m pytest tests -x -q -m "not gpu"

test-gpu: all
	python -m pytest tests -x -q -m gpu

clean:
	$(MAKE) -C starrynight_b200/csrc clean
	$(MAKE) -C oracle clean
	rm -f driver/starrynight-b200

.PHONY: all test-cpu test-gpu clean
