MBC_LIB_PATH=$PWD/minibase-columnar-database_b200/csrc/variants/libmbcol_prof.so timeout 120 python scripts/profile_scan.py 100000000 2 2>&1 | tail -24
