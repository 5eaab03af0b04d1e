# Convenience targets; the authoritative build entry point is __graft_entry__.build().
PY ?= python

build:
	$(PY) -c "import __graft_entry__ as g; g.build()"

test:            ## CPU suite (no GPU needed)
	$(PY) -m pytest tests -q -m "not gpu"

test-gpu:        ## on a B200
	$(PY) -m pytest tests -q -m gpu

bench:
	$(PY) bench.py

hostsim-campaign:
	bash tools/hostsim_campaign.sh

clean:
	$(MAKE) -C acg_b200/csrc clean
	$(MAKE) -C oracle clean || true
	$(MAKE) -C examples clean
	$(MAKE) -C tests/hostsim clean

.PHONY: build test test-gpu bench hostsim-campaign clean
