#!/bin/bash
# final binary: smoke + the ABI / proof / witness parity tests + one short bench line
timeout 600 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 900 python -m pytest tests/test_abi.py tests/test_gpu_prove.py tests/test_gpu_witness.py tests/test_gpu_ntt.py -m gpu -x -q 2>&1 | tail -2
timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extra 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('value %.1f e2e %.1f wit %.0f'%(d['value'],d['e2e']['value'],d['witness']['value']))"
