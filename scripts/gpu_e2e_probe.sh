for d in 32 8 1; do MBC_LATE_DIV=$d timeout 200 python scripts/e2e_probe.py 100000000 0.01,0.03,0.1,0.5 2>&1 | tail -4; done
