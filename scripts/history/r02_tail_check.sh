#!/bin/bash
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 1
timeout 150 python -m pytest tests/test_ldl_gpu.py "tests/test_configs_gpu.py::test_c4_full_size" "tests/test_configs_gpu.py::test_c2_full_size" -x -q -m gpu 2>&1 | tail -n 1
