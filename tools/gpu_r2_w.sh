#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_prove.py -m gpu -x -q --durations=4 2>&1 | tail -9
