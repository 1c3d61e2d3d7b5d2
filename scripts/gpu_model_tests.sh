#!/bin/bash
# The inertial-set (tcmp_model) tests only.
python -m pytest tests/test_gpu_parity.py tests/test_gpu_dropin.py -x -q -k "model or inertial" 2>&1 | tail -15
