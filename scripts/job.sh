#!/bin/bash
# scratch GPU job: deferred / abandoned encodes; API tests (interrupt + resume among them)
python -m pytest tests/test_gpu_png.py tests/test_gpu_api.py -m gpu -x -q 2>&1 | tail -15
