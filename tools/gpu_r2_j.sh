#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/dbg_config3.py 2>&1 | tail -14
