for L in 200 800 1600 3200 6400 12800; do timeout 120 python scripts/ldl_tune.py one $L 0 2>&1 | grep refactor_ms; done
