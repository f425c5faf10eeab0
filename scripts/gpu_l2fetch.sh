for g in 32 64 128; do echo "== L2 fetch $g"; MBC_L2_FETCH=$g timeout 120 python scripts/profile_scan.py 100000000 3 2>&1 | tail -4; done
echo "== default"; timeout 120 python scripts/profile_scan.py 100000000 3 2>&1 | tail -3
