timeout 300 python -m pytest tests/test_bench_contract.py -m gpu -x -q 2>&1 | tail -3
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
VARIANTS="qb16m4 qb8m3 qb8m4" WL="MultiviewC" bash scripts/list_timing.sh 2>&1 | grep -E "^default|^qb" 
for v in qb16m4 qb8m3; do VFA_B200_LIB=$PWD/build/variants/libvfa_$v.so timeout 100 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:qlist_build -c 2 python scripts/quick_time.py MultiviewC 4 0 2>&1 | grep -E "gpu__time_duration" | tail -1; done
