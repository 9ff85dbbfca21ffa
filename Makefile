# Developer entry points (the reference's pre-commit hook runs `cargo fmt --check` + `cargo test`; this is our equivalent).
.PHONY: build test test-gpu bench clean
build:
	python -c "import __graft_entry__ as g; g.build()"
test: build
	python -m pytest tests -x -q -m "not gpu"
test-gpu: build
	python -m pytest tests -x -q -m gpu
bench: build
	python bench.py --gpus 1 --steps 10 --warmup 3
clean:
	$(MAKE) -C rust-raytracer_b200 clean
	$(MAKE) -C oracle clean
